"""time the contract step (graph replay, L2 flushed) and its halves for the library in UGL_LIB_PATH"""
import sys, os, json
sys.path.insert(0, ".")
import torch
import bench
dev = torch.device("cuda:0")
B, H, W, L = 8, 256, 832, 4
import argparse
from unsupervised_depth_opticalflow_egomotion_b200 import ops
from unsupervised_depth_opticalflow_egomotion_b200.synth import make_triplet
MODE = os.environ.get("FLOW_MODE", "noise")
t = make_triplet(B, H, W, L, 4 if MODE == "rigid" else 1, seed=1234, flow_px=10.0, flow_mode=MODE, device=dev)
print("flow stats", MODE, [float(f.abs().mean()) for f in t.flows_fwd], "dx", float((t.flows_fwd[0][..., 1:] - t.flows_fwd[0][..., :-1]).abs().mean()))
pl, pc, pr = ops.image_pyramids([t.img_l, t.img, t.img_r], L, ["box"] * 3) if False else [ops.image_pyramid(x, L, "box") for x in (t.img_l, t.img, t.img_r)]
gl = torch.tensor([0.15, 0.85, 10.0, 0.01], device=dev).view(4, 1).repeat(1, B) / B
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def timeit(phase, n=20, mode="single_pass"):
    out = ops.flow_loss_step(pl, pc, pr, t.flows_fwd, t.flows_bwd, gl, L, phase="both", mode=mode)
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for _ in range(3): ops.flow_loss_step(pl, pc, pr, t.flows_fwd, t.flows_bwd, gl, L, out=out, phase=phase, mode=mode)
        torch.cuda.synchronize()
        with torch.cuda.graph(g, stream=s):
            ops.flow_loss_step(pl, pc, pr, t.flows_fwd, t.flows_bwd, gl, L, out=out, phase=phase, mode=mode)
    tot = 0.0
    for _ in range(n + 3):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
        if _ >= 3: tot += e0.elapsed_time(e1)
    return tot / n
print("fused_step %.4f ms" % timeit("both", mode="fused_step"), flush=True)
print(MODE, os.environ.get("UGL_LIB_PATH", "default"), "both %.4f fwd %.4f bwd %.4f ms" % (timeit("both"), timeit("forward"), timeit("backward")), flush=True)

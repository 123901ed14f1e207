"""does running the step on batch slices in concurrent streams (photo of one slice under the stencil of another) beat one launch?"""
import sys, os
sys.path.insert(0, ".")
import torch
from unsupervised_depth_opticalflow_egomotion_b200 import ops
from unsupervised_depth_opticalflow_egomotion_b200.synth import make_triplet
dev = torch.device("cuda:0")
B, H, W, L = 8, 256, 832, 4
t = make_triplet(B, H, W, L, 1, seed=1234, flow_px=10.0, device=dev)
pl, pc, pr = [ops.image_pyramid(x, L, "box") for x in (t.img_l, t.img, t.img_r)]
gl = torch.tensor([0.15, 0.85, 10.0, 0.01], device=dev).view(4, 1).repeat(1, B) / B
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

def build(nsl, stagger):
    sl = [slice(k * B // nsl, (k + 1) * B // nsl) for k in range(nsl)]
    ins = [([x[s].contiguous() for x in pl], [x[s].contiguous() for x in pc], [x[s].contiguous() for x in pr],
            [x[s].contiguous() for x in t.flows_fwd], [x[s].contiguous() for x in t.flows_bwd], gl[:, s].contiguous()) for s in sl]
    outs = [ops.flow_loss_step(*i[:5], i[5], L) for i in ins]
    streams = [torch.cuda.Stream() for _ in range(nsl)]
    main = torch.cuda.Stream()
    g = torch.cuda.CUDAGraph()
    torch.cuda.synchronize()
    with torch.cuda.stream(main):
        with torch.cuda.graph(g, stream=main):
            for k in range(nsl):
                streams[k].wait_stream(main)
                with torch.cuda.stream(streams[k]):
                    ops.flow_loss_step(*ins[k][:5], ins[k][5], L, out=outs[k])
            for k in range(nsl):
                main.wait_stream(streams[k])
    return g

def timeit(g, n=20):
    tot = 0.0
    for it in range(n + 3):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
        if it >= 3: tot += e0.elapsed_time(e1)
    return tot / n

for nsl in (1, 2, 4, 8):
    g = build(nsl, False)
    print("slices", nsl, "ms %.4f" % timeit(g), flush=True)

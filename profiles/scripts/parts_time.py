"""Contract step (B=8, 256x832, 4 levels) for the library in UGL_LIB_PATH: graph-replay time of the whole step and of each launch alone
(L2 flushed before each), plus a checksum of the outputs (sha1 of losses + flow gradients) so variants can be compared for bit identity.
One JSON line."""
import sys, os, json, hashlib
sys.path.insert(0, ".")
import torch
from unsupervised_depth_opticalflow_egomotion_b200 import ops
from unsupervised_depth_opticalflow_egomotion_b200.synth import make_triplet

dev = torch.device("cuda:0")
B, H, W, L = 8, 256, 832, 4
N = int(os.environ.get("N", "20"))
MODE = os.environ.get("FLOW_MODE", "noise")
t = make_triplet(B, H, W, L, 4 if MODE == "rigid" else 1, seed=1234, flow_px=10.0, flow_mode=MODE, device=dev)
pl, pc, pr = [ops.image_pyramid(x, L, "box") for x in (t.img_l, t.img, t.img_r)]
gl = torch.tensor([0.15, 0.85, 10.0, 0.01], device=dev).view(4, 1).repeat(1, B) / B
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
out = ops.flow_loss_step(pl, pc, pr, t.flows_fwd, t.flows_bwd, gl, L)
torch.cuda.synchronize()
h = hashlib.sha1()
for x in [out["loss"]] + out["gf"] + out["gb"]:
    h.update(x.detach().cpu().numpy().tobytes())


def timeit(phase):
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for _ in range(2):
            ops.flow_loss_step(pl, pc, pr, t.flows_fwd, t.flows_bwd, gl, L, out=out, phase=phase)
        torch.cuda.synchronize()
        with torch.cuda.graph(g, stream=s):
            ops.flow_loss_step(pl, pc, pr, t.flows_fwd, t.flows_bwd, gl, L, out=out, phase=phase)
    ts = []
    for k in range(N + 3):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
        if k >= 3:
            ts.append(e0.elapsed_time(e1))
    ts.sort()
    return round(1e3 * sum(ts) / len(ts), 1), round(1e3 * ts[len(ts) // 2], 1)


def sha_out():
    torch.cuda.synchronize()
    hh = hashlib.sha1()
    for x in [out["loss"]] + out["gf"] + out["gb"]:
        hh.update(x.detach().cpu().numpy().tobytes())
    return hh.hexdigest()[:12]


res = {"lib": os.path.basename(os.environ.get("UGL_LIB_PATH", "default")), "sha": h.hexdigest()[:12], "loss": float(out["loss"].sum())}
for ph in os.environ.get("PHASES", "both,photo,norm,stencil,finalize").split(","):
    res[ph + "_us(mean,med)"] = timeit(ph)
# stale-data check for the programmatic dependent launches: one graph, inputs changed in place between replays
g2 = torch.cuda.CUDAGraph()
s2 = torch.cuda.Stream()
with torch.cuda.stream(s2):
    ops.flow_loss_step(pl, pc, pr, t.flows_fwd, t.flows_bwd, gl, L, out=out)
    torch.cuda.synchronize()
    with torch.cuda.graph(g2, stream=s2):
        ops.flow_loss_step(pl, pc, pr, t.flows_fwd, t.flows_bwd, gl, L, out=out)
shas = []
for k in range(3):
    for f in t.flows_fwd + t.flows_bwd:
        f.mul_(0.9)
    for x in pc:
        x.mul_(0.95)
    torch.cuda.synchronize()
    g2.replay()
    shas.append(sha_out())
res["replay_shas"] = shas
print(json.dumps(res), flush=True)

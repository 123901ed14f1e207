"""fused_step vs single_pass (and timing)"""
import sys, os
sys.path.insert(0, ".")
import torch
from unsupervised_depth_opticalflow_egomotion_b200 import ops
from unsupervised_depth_opticalflow_egomotion_b200.synth import make_triplet
dev = torch.device("cuda:0")
def cmp(B, H, W, L):
    t = make_triplet(B, H, W, L, 1, seed=5, flow_px=6.0, device=dev)
    pl, pc, pr = [ops.image_pyramid(x, L, "box") for x in (t.img_l, t.img, t.img_r)]
    gl = (torch.tensor([0.15, 0.85, 10.0, 0.01], device=dev).view(4, 1) * (1 + torch.arange(B, device=dev).float().view(1, B))) / B
    a = ops.flow_loss_step(pl, pc, pr, t.flows_fwd, t.flows_bwd, gl, L, mode="single_pass")
    b = ops.flow_loss_step(pl, pc, pr, t.flows_fwd, t.flows_bwd, gl, L, mode="fused_step")
    torch.cuda.synchronize()
    le = ((a["loss"] - b["loss"]).abs() / a["loss"].abs()).max().item()
    ge = max(((x - y).abs().max() / x.abs().max()).item() for x, y in zip(a["gf"] + a["gb"], b["gf"] + b["gb"]))
    print((B, H, W, L), "loss rel %.2e grad rel %.2e" % (le, ge), flush=True)
for cfg in [(2, 64, 208, 4), (1, 36, 52, 2), (2, 96, 160, 4), (1, 256, 832, 4), (3, 39, 57, 1)]:
    cmp(*cfg)

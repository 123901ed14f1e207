import sys, os
sys.path.insert(0, ".")
import torch
from unsupervised_depth_opticalflow_egomotion_b200 import ops
from unsupervised_depth_opticalflow_egomotion_b200.synth import make_triplet
dev = torch.device("cuda:0")
B, H, W, L = 8, 256, 832, 4
mode = sys.argv[1] if len(sys.argv) > 1 else "fused_step"
t = make_triplet(B, H, W, L, 1, seed=1234, flow_px=10.0, device=dev)
pl, pc, pr = [ops.image_pyramid(x, L, "box") for x in (t.img_l, t.img, t.img_r)]
gl = torch.tensor([0.15, 0.85, 10.0, 0.01], device=dev).view(4, 1).repeat(1, B) / B
out = None
for _ in range(4):
    out = ops.flow_loss_step(pl, pc, pr, t.flows_fwd, t.flows_bwd, gl, L, out=out, mode=mode)
torch.cuda.synchronize()
print(out["loss"].sum().item())

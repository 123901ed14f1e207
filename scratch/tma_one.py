import sys, torch
sys.path.insert(0, ".")
from unsupervised_depth_opticalflow_egomotion_b200 import ops
from unsupervised_depth_opticalflow_egomotion_b200.synth import make_triplet
dev = torch.device("cuda:0")
B, H, W, L = 1, 64, 208, 3
ops.SINGLE_PASS_VARIANT = sys.argv[1] if len(sys.argv) > 1 else "split_tma"
t = make_triplet(B, H, W, L, 1, seed=3, flow_px=5.0, device=dev)
pl, pc, pr = (ops.image_pyramid(x, L, "box") for x in (t.img_l, t.img, t.img_r))
ff = [f.clone().requires_grad_(True) for f in t.flows_fwd]
fb = [f.clone().requires_grad_(True) for f in t.flows_bwd]
loss = ops.flow_loss(pl, pc, pr, ff, fb, L, as_matrix=True)
torch.cuda.synchronize()
print(loss)

import sys
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import torch
from oracle import loss_port as P
from unsupervised_depth_opticalflow_egomotion_b200 import losses, ops
from util import load_golden, golden_triplet
dev = torch.device("cuda:0")
d = load_golden("geom_mode_s3_mt")
t = golden_triplet(d)
lf = lambda xs: [x.detach().to(dev).requires_grad_(True) for x in xs]
ff, fb = lf(t.flows_fwd), lf(t.flows_bwd)
disp, disp_l, disp_r = lf(t.disp), lf(t.disp_l), lf(t.disp_r)
pose = t.pose.to(dev).requires_grad_(True)
loss, masks = losses.GeometryLoss(3).forward_losses(t.img_l.to(dev), t.img.to(dev), t.img_r.to(dev), ff, fb, disp, disp_l, disp_r, pose, t.K.to(dev), t.K_inv.to(dev))
live = [k for k in loss if "out_" + k in d and d["out_" + k].numel() == 1 and loss[k].requires_grad]
for k in live:
    print(k, float(loss[k]), float(d["out_" + k]), abs(float(loss[k]) - float(d["out_" + k])) / max(abs(float(d["out_" + k])), 1e-30))
# per-term gradient wrt flows_bwd_0
for k in live:
    g = torch.autograd.grad(P.GEOM_WEIGHTS[k] * loss[k].mean(), fb[0], retain_graph=True, allow_unused=True)[0]
    print(k, None if g is None else float(g.abs().max()))
tot = sum(P.GEOM_WEIGHTS[k] * loss[k].mean() for k in live)
g = torch.autograd.grad(tot, fb[0])[0].cpu()
ref = d["grad_flows_bwd_0"]
diff = (g - ref).abs()
scale = ref.abs().max()
bad = (diff > 1.5e-4 * scale).nonzero()
print("scale", float(scale), "bad", bad.tolist())
rf = P.rigid_flow(t.disp[0], t.pose[:, 0], t.K)
for b, c, i, j in bad.tolist():
    print((b, c, i, j), "got %.4e ref %.4e" % (g[b, c, i, j], ref[b, c, i, j]), "rf-f:", (rf[b, :, i, j] - t.flows_bwd[0][b, :, i, j]).tolist(),
          "bwd_mask-ish dyn", float(d["aux_dyn_b_0"][b, 0, i, j]), "occ", float(d["aux_occ_b_0"][b, 0, i, j]), "valid", float(d["aux_valid_b_0"][b, 0, i, j]))

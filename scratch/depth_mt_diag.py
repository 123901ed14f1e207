import sys
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import torch
from oracle import loss_port as P
from unsupervised_depth_opticalflow_egomotion_b200 import losses, ops
from util import load_golden, golden_triplet
dev = torch.device("cuda:0")
d = load_golden("depth_mode_texture_mt")
t = golden_triplet(d)
lf = lambda xs: [x.detach().to(dev).requires_grad_(True) for x in xs]
res = {}
for fused in (True, False):
    disp, disp_l, disp_r = lf(t.disp), lf(t.disp_l), lf(t.disp_r)
    pose = t.pose.to(dev).requires_grad_(True)
    loss, masks = losses.DepthLoss(3, "texture").forward_losses(t.img_l.to(dev), t.img.to(dev), t.img_r.to(dev), disp, disp_l, disp_r, pose, t.K.to(dev), fused=fused)
    per = {}
    for k in ("loss_depth_pixel", "loss_depth_ssim", "loss_depth_consis", "loss_depth_smooth"):
        g = torch.autograd.grad(P.GEOM_WEIGHTS[k] * loss[k].mean(), disp[0], retain_graph=True, allow_unused=True)[0]
        per[k] = None if g is None else g.cpu()
    res[fused] = per
# oracle per term fp32 / fp64
def oracle(dtype):
    cv = lambda x: x.detach().to(dtype)
    disp = [cv(x).requires_grad_(True) for x in t.disp]
    loss = P.depth_mode_loss(cv(t.img_l), cv(t.img), cv(t.img_r), disp, [cv(x) for x in t.disp_l], [cv(x) for x in t.disp_r], cv(t.pose), cv(t.K), 3, "texture")
    return {k: torch.autograd.grad(P.GEOM_WEIGHTS[k] * loss[k].mean(), disp[0], retain_graph=True)[0] for k in ("loss_depth_pixel", "loss_depth_ssim", "loss_depth_consis", "loss_depth_smooth")}
o32, o64 = oracle(torch.float32), oracle(torch.float64)
tot_ref = d["grad_disp_0"]
scale = float(tot_ref.abs().max())
print("total scale", scale)
for k in o32:
    a, b, c, e = res[True][k], res[False][k], o32[k], o64[k]
    bad = ((a - c).abs() > 1e-4 * scale).nonzero().tolist()
    print(k, "fused-vs-o32 %.3e composed-vs-o32 %.3e o32-vs-o64 %.3e fused-vs-o64 %.3e" % (float((a - c).abs().max()) / scale, float((b - c).abs().max()) / scale,
          float((c - e).abs().max()) / scale, float((a - e.float()).abs().max()) / scale), "bad px", [(x[2], x[3]) for x in bad][:12])

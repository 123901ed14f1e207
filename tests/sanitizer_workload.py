import sys, torch
sys.path.insert(0, ".")
from unsupervised_depth_opticalflow_egomotion_b200 import ops, losses
from unsupervised_depth_opticalflow_egomotion_b200.synth import make_triplet
dev = torch.device("cuda:0")
t = make_triplet(2, 40, 72, 4, 3, seed=3, flow_mode="rigid", oob_fraction=0.1).to(dev)
pl, pc, pr = (ops.image_pyramid(x, 4, "box") for x in (t.img_l, t.img, t.img_r))
for mode, variant in (("single_pass", "split"), ("single_pass", "split_plain"), ("single_pass", "fused"), ("recompute", "split")):
    ops.SINGLE_PASS_VARIANT = variant     # split: photometry kernel + stencil kernel (TMA staging where the widths allow: levels 0..2 of 72 px)
    ff = [f.detach().requires_grad_(True) for f in t.flows_fwd]; fb = [f.detach().requires_grad_(True) for f in t.flows_bwd]
    loss = ops.flow_loss(pl, pc, pr, ff, fb, 4, as_matrix=True, mode=mode)
    g = torch.autograd.grad(loss.sum(), ff + fb)
ops.SINGLE_PASS_VARIANT = "split"
gl = torch.full((4, 2), 0.5, device=dev)
for step_mode in ("fused_step", "single_pass"):      # ugl_flow_loss_step (TMA-staged stencil writing the gradients) and forward_grad + combine
    ops.flow_loss_step(pl, pc, pr, t.flows_fwd, t.flows_bwd, gl, 4, mode=step_mode)
t2 = make_triplet(1, 64, 96, 3, 1, seed=4).to(dev)   # every level a multiple of 4 wide: all levels TMA-staged
p2 = [ops.image_pyramid(x, 3, "box") for x in (t2.img_l, t2.img, t2.img_r)]
ops.flow_loss_step(p2[0], p2[1], p2[2], t2.flows_fwd, t2.flows_bwd, torch.full((4, 1), 0.5, device=dev), 3)
leaves = [x.detach().requires_grad_(True) for x in t.flows_fwd + t.flows_bwd + t.disp + t.disp_l + t.disp_r + [t.pose]]
ff, fb, d, dl, dr, pose = leaves[0:4], leaves[4:8], leaves[8:11], leaves[11:14], leaves[14:17], leaves[17]
loss, _ = losses.GeometryLoss(3).forward_losses(t.img_l, t.img, t.img_r, ff, fb, d, dl, dr, pose, t.K, t.K_inv)
sum(v.mean() for v in loss.values()).backward()
for variant in ("texture", "ssim", "live"):
    for fused in (True, "ops", False):
        loss, _ = losses.DepthLoss(3, variant).forward_losses(t.img_l, t.img, t.img_r, d, dl, dr, pose, t.K, fused=fused)
        sum(v.mean() for v in loss.values()).backward()
loss, masks = losses.GeometryLoss(3).forward_losses(t.img_l, t.img, t.img_r, ff, fb, d, dl, dr, pose, t.K, t.K_inv, fused=False)
sum(v.mean() for v in loss.values()).backward()
loss, masks = losses.GeometryLoss(3).forward_losses(t.img_l, t.img, t.img_r, ff, fb, d, dl, dr, pose, t.K, t.K_inv)
_ = masks["occ_b"], masks["dist_f"], masks["rigid_f"]
losses.total_loss(loss, {k: 1.0 for k in loss}).backward()
ops.image_pyramids((t.img, t.img_l, t.img_r), 4, ("box", ("bilinear", "area"), "bilinear"))
ops.forward_splat(torch.ones(2, 1, 40, 72, device=dev), t.flows_fwd[0].detach(), True)
x = torch.rand(1, 8, 20, 30, device=dev, requires_grad=True); fl = (3 * torch.randn(1, 2, 20, 30, device=dev)).requires_grad_(True)
ops.warp_flow(x, fl, True).sum().backward()
torch.cuda.synchronize(); print("sanitizer workload done")
assert int(ops.selftest_packed_pairs(dev, blocks=2, windows_per_thread=4).sum()) == 0
ops.frames_from_u8([(255 * torch.rand(2, 3, 40, 72, device=dev)).to(torch.uint8) for _ in range(3)])
ops.DEPTH_PHOTO_SINGLE_PASS = False
loss, _ = losses.DepthLoss(3, "live").forward_losses(t.img_l, t.img, t.img_r, d, dl, dr, pose, t.K)
sum(v.mean() for v in loss.values()).backward()
ops.DEPTH_PHOTO_SINGLE_PASS = True
torch.cuda.synchronize(); print("sanitizer workload done (incl. packed-pair self-test)")

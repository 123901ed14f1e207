"""GeometryLoss with and without step_weights (fused geom training step) on random shapes: losses bit-identical, gradients to 2e-6."""
import sys, random
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import torch
from oracle import loss_port as P           # weights table only
from unsupervised_depth_opticalflow_egomotion_b200 import losses
from unsupervised_depth_opticalflow_egomotion_b200.synth import make_triplet
from util import rel_err
dev = torch.device("cuda:0")
random.seed(11)
bad = 0
leaf = lambda xs: [x.detach().clone().requires_grad_(True) for x in xs]
for it in range(8):
    H, W, B = 4 * random.randint(6, 40), 4 * random.randint(8, 70), random.choice([1, 2, 3])
    t = make_triplet(B, H, W, 4, 3, seed=200 + it, flow_mode="rigid").to(dev)
    res = {}
    for step in (False, True):
        ff, fb = leaf(t.flows_fwd), leaf(t.flows_bwd)
        disp, disp_l, disp_r = leaf(t.disp), leaf(t.disp_l), leaf(t.disp_r)
        pose = t.pose.detach().clone().requires_grad_(True)
        loss, masks = losses.GeometryLoss(3).forward_losses(t.img_l, t.img, t.img_r, ff, fb, disp, disp_l, disp_r, pose, t.K, t.K_inv,
                                                            step_weights=P.GEOM_WEIGHTS if step else None)
        total = losses.total_loss(loss, P.GEOM_WEIGHTS)
        g = torch.autograd.grad(total, ff[:3] + fb[:3] + disp + disp_l + disp_r + [pose])
        res[step] = (loss, total, g)
    a, b = res[True], res[False]
    same_loss = all(torch.equal(a[0][k].detach(), b[0][k].detach()) for k in b[0])
    e = max(float(rel_err(x, y)) for x, y in zip(a[2], b[2]))
    ok = same_loss and e < 2e-6 and all(torch.isfinite(x).all() for x in a[2])
    bad += (not ok)
    print(it, (B, H, W), "OK" if ok else "BAD", "losses equal" if same_loss else "LOSSES DIFFER", "max grad rel err %.1e" % e, flush=True)
print("bad", bad)

import sys, random
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import torch
from unsupervised_depth_opticalflow_egomotion_b200 import ops
from unsupervised_depth_opticalflow_egomotion_b200.synth import make_triplet
from util import rel_err, loss_rel_err
dev = torch.device("cuda:0")
random.seed(7)
bad = 0
for it in range(14):
    L = random.choice([1, 2, 3, 4])
    even = random.random() < 0.6
    if even:
        m = 1 << (L - 1)
        H, W = m * random.randint(3, 40), m * random.randint(3, 60)
    else:
        L = 1
        H, W = random.randint(3, 150), random.randint(3, 200)
    B = random.choice([1, 2, 3])
    scales = random.randint(1, L)
    t = make_triplet(B, H, W, L, 1, seed=100 + it, flow_px=random.choice([1.0, 4.0, 15.0]), oob_fraction=0.05).to(dev)
    pl, pc, pr = ([t.img_l], [t.img], [t.img_r]) if L == 1 else [ops.image_pyramid(x, L, "box") for x in (t.img_l, t.img, t.img_r)]
    gl = (torch.rand(4, B) + 0.5).to(dev)
    st = ops.flow_loss_step(pl, pc, pr, t.flows_fwd, t.flows_bwd, gl, scales)
    a_ff = [f.detach().clone().requires_grad_(True) for f in t.flows_fwd]
    a_fb = [f.detach().clone().requires_grad_(True) for f in t.flows_bwd]
    errs = []
    for variant in ("split", "fused", "split_plain"):
        ops.SINGLE_PASS_VARIANT = variant
        al = ops.flow_loss(pl, pc, pr, a_ff, a_fb, scales, as_matrix=True)
        ag = torch.autograd.grad(al, a_ff[:scales] + a_fb[:scales], grad_outputs=gl)
        e_l = float(loss_rel_err(st["loss"], al))
        e_g = max(float(rel_err(a, b)) for a, b in zip(st["gf"] + st["gb"], ag))
        errs.append((variant, e_l, e_g))
    ops.SINGLE_PASS_VARIANT = "split"
    ok = all(e[1] < 2e-6 and e[2] < 5e-6 for e in errs) and all(torch.isfinite(x).all() for x in st["gf"] + st["gb"])
    bad += (not ok)
    print(it, (B, H, W, L, scales), "OK" if ok else "BAD", [(v, "%.1e" % a, "%.1e" % b) for v, a, b in errs], flush=True)
print("bad", bad)

"""GPU check of the split single-pass variants against the fused kernel (and timing)."""
import sys, time
import torch
sys.path.insert(0, ".")
from unsupervised_depth_opticalflow_egomotion_b200 import ops
from unsupervised_depth_opticalflow_egomotion_b200.synth import make_triplet

dev = torch.device("cuda:0")

def run(B, H, W, L, variant, seed=3, flow_px=5.0):
    ops.SINGLE_PASS_VARIANT = variant
    t = make_triplet(B, H, W, L, 1, seed=seed, flow_px=flow_px, device=dev)
    pl, pc, pr = (ops.image_pyramid(x, L, "box") for x in (t.img_l, t.img, t.img_r))
    ff = [f.clone().requires_grad_(True) for f in t.flows_fwd]
    fb = [f.clone().requires_grad_(True) for f in t.flows_bwd]
    loss = ops.flow_loss(pl, pc, pr, ff, fb, L, as_matrix=True)
    w = torch.tensor([0.15, 0.85, 10.0, 0.01], device=dev)
    tot = (loss.mean(1) * w).sum()
    g = torch.autograd.grad(tot, ff + fb)
    torch.cuda.synchronize()
    return loss.detach().cpu(), [x.detach().cpu() for x in g]

for (B, H, W, L) in [(2, 64, 208, 4), (1, 36, 52, 2), (2, 64, 208, 3), (2, 96, 160, 4), (1, 256, 832, 4)]:
    ref_l, ref_g = run(B, H, W, L, "fused")
    for v in ("split_plain", "split_tma", "split"):
        try:
            l, g = run(B, H, W, L, v)
        except Exception as e:
            print((B, H, W, L), v, "ERR", str(e)[:120]); continue
        le = ((l - ref_l).abs() / ref_l.abs().clamp_min(1e-30)).max().item()
        ge = max(((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item() for a, b in zip(g, ref_g))
        print((B, H, W, L), v, "loss rel %.2e grad rel %.2e" % (le, ge), flush=True)

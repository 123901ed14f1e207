"""Time warp_flow's grad_x (the deterministic scatter) in both forms at the flow network's feature-warp shapes (pwc_tf.py:94-95).
CUDA events around the backward of the op alone (grad_x only), L2 not flushed (the op is called back to back in training too)."""
import json, sys, torch
sys.path.insert(0, ".")
from unsupervised_depth_opticalflow_egomotion_b200 import ops

dev = torch.device("cuda:0")
shapes = [(8, 3, 256, 832, 4.0), (8, 16, 128, 416, 2.0), (8, 32, 64, 208, 1.0), (8, 64, 32, 104, 0.5), (8, 96, 16, 52, 0.25),
          (8, 32, 64, 208, 12.0)]
rows = []
for B, C, H, W, px in shapes:
    g = torch.Generator().manual_seed(1)
    x = torch.rand(B, C, H, W, generator=g).to(dev).requires_grad_(True)
    flow = (px * torch.randn(B, 2, H, W, generator=g)).to(dev)
    go = torch.randn(B, C, H, W, generator=g).to(dev)
    row = {"shape": [B, C, H, W], "flow_px": px}
    ref = None
    for form in ("global", "tile_local"):
        ops.SCATTER_FORM = form
        out = ops.warp_flow(x, flow, False)
        for _ in range(3):
            gx, = torch.autograd.grad(out, [x], go, retain_graph=True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(20):
            gx, = torch.autograd.grad(out, [x], go, retain_graph=True)
        e1.record()
        torch.cuda.synchronize()
        row[form + "_us"] = round(e0.elapsed_time(e1) * 1000 / 20, 1)
        ref = gx if ref is None else ref
        row["same_bits"] = bool(torch.equal(ref, gx))
    rows.append(row)
    print(json.dumps(row), flush=True)

import sys, os
sys.path.insert(0, ".")
import torch, bench
dev = torch.device("cuda:0")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for wl in sys.argv[1:]:
    r = bench.mode_step(wl, 8, 256, 832, dev, 20, 3, flush, 0, graph=True)
    print(wl, "ms %.4f launches %d frac %.4f" % (r["ms_per_step"], r["launches_per_step"], r["hbm_frac"]), flush=True)

import sys, os
sys.path.insert(0, ".")
import torch, bench
dev = torch.device("cuda:0")
wl = sys.argv[1]
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
r = bench.mode_step(wl, 8, 256, 832, dev, 3, 2, flush, 0, graph=False)
print(wl, r["ms_per_step"], r["launches_per_step"])

"""First-call anomaly of torch CPU math on some boxes (oracle side): fresh processes, with / without a single-threaded warm-up."""
import sys, os, subprocess
child = r'''
import sys, torch
sys.path.insert(0, "."); sys.path.insert(0, "tests")
mode = sys.argv[1]
if mode == "warm":
    n = torch.get_num_threads(); torch.set_num_threads(1)
    x = torch.linspace(-8, 1, 64).reshape(1, 2, 4, 8)
    torch.softmax(x, 1); torch.exp(x); (x ** 2); torch.sqrt(x.abs()); torch.log(x.abs() + 1)
    torch.set_num_threads(n)
from oracle import loss_port as P
from unsupervised_depth_opticalflow_egomotion_b200.synth import make_triplet
t = make_triplet(2, 64, 208, 1, 1, seed=31, flow_px=5.0, oob_fraction=0.1)
from_l = P.flow_backwarp(t.img_l, t.flows_bwd[0], True)
from_r = P.flow_backwarp(t.img_r, t.flows_fwd[0], True)
o = P.occlusion_weights([from_l], [t.img], [from_r], 1, soft=True)
o64 = P.occlusion_weights([from_l.double()], [t.img.double()], [from_r.double()], 1, soft=True)
e = float((o["w_bwd"][0].double() - o64["w_bwd"][0]).abs().max())
d = torch.cat([o["diff_bwd"][0], o["diff_fwd"][0]], 1)
print(mode, "first-call cpu32 vs cpu64 abs err %.3e" % e)
'''
open("/tmp/child.py", "w").write(child)
def run(mode, n, env=None):
    errs = []
    e2 = dict(os.environ); e2.update(env or {})
    for i in range(n):
        out = subprocess.run([sys.executable, "/tmp/child.py", mode], capture_output=True, text=True, env=e2).stdout.strip().split()
        errs.append(float(out[-1]) if out else -1.0)
    bad = sum(e > 1e-5 for e in errs)
    print(mode, env or "", "bad %d of %d" % (bad, len(errs)), sorted(set("%.1e" % e for e in errs)), flush=True)
    return bad

n = int(sys.argv[1]) if len(sys.argv) > 1 else 12
print(subprocess.run("lscpu | grep -E 'Model name|Hypervisor|L3'", shell=True, capture_output=True, text=True).stdout)
if run("cold", n):
    run("cold", n)
    run("warm", n)
    run("cold", n, {"ATEN_CPU_CAPABILITY": "avx2", "MKL_ENABLE_INSTRUCTIONS": "AVX2"})
    run("cold", n, {"OMP_NUM_THREADS": "1", "MKL_NUM_THREADS": "1"})
    run("cold", n, {"MKL_CBWR": "COMPATIBLE"})

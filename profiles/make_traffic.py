"""Turn one `ncu --set full` capture of the step's kernels into profiles/traffic.json (read by bench.py for `roofline.traffic`,
the issue fraction and the per-kernel counters).
usage: ncu -i X.ncu-rep --page raw --csv | python profiles/make_traffic.py CAPTURE_NAME > profiles/traffic.json"""
import csv, json, sys
rows = list(csv.reader(sys.stdin))
h, units = rows[0], rows[1]
col = {n: i for i, n in enumerate(h)}
def val(r, name, scale_units=True):
    v = float(r[col[name]].replace(",", "") or 0)
    u = units[col[name]]
    if scale_units:
        v *= {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1e3, "us": 1.0, "ns": 1e-3, "second": 1e6}.get(u, 1.0)
    return v
out = {"capture": sys.argv[1] if len(sys.argv) > 1 else "?", "kernels": {}}
for r in rows[2:]:
    name = r[col["Kernel Name"]]
    short = name.split("<")[0].split("(")[0].replace("void ", "").replace("ugl::", "").strip()
    out["kernels"][short] = {
        "kernel": name[:120],
        "dram_bytes_read": val(r, "dram__bytes_read.sum"), "dram_bytes_write": val(r, "dram__bytes_write.sum"),
        "time_us": val(r, "gpu__time_duration.sum"), "warp_instructions": val(r, "smsp__inst_executed.sum", False),
        "shared_wavefronts": val(r, "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", False),
        "issue_active_pct": val(r, "smsp__issue_active.avg.pct_of_peak_sustained_active", False),
        "warps_active_pct": val(r, "sm__warps_active.avg.pct_of_peak_sustained_active", False),
        "fma_pipe_pct": val(r, "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", False),
        "registers": val(r, "launch__registers_per_thread", False),
    }
json.dump(out, sys.stdout, indent=1)

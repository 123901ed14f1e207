import json,sys
d=json.load(open(sys.argv[1])); r=d["roofline"]; print(sys.argv[1], "ms/step", round(d["ms_per_step"],4), "kernel_ms", round(r["kernel_ms"],4), "combine", round(r["combine_ms"],4), "frac", round(r["frac"],4), d["clocks"]["sm_mhz"])

"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list into per-kernel totals of ONE step.

usage: python profiles/summarise_launches.py launches.csv MARKER
A step is the span between the last two launches whose kernel name contains MARKER (the first kernel of a step).
Durations are ncu's serialised cold-cache times: the SHARES are what carries over to the timed run, not the absolutes."""
import collections
import csv
import sys


def main(path, marker, exclude=()):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    hdr, data = rows[0], rows[1:]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    idx = [i for i, r in enumerate(data) if marker in r[ki]]
    step = [r for r in data[idx[-2]:idx[-1]] if not any(x in r[ki] for x in exclude)]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in step:
        v = float(r[vi].replace(",", "")) / (1000.0 if r[ui] == "ns" else 1.0)
        agg[r[ki][:110]][0] += 1
        agg[r[ki][:110]][1] += v
    tot = sum(v[1] for v in agg.values())
    print("%s: %d launches per step, %.1f us (serialised, cold cache)" % (path, len(step), tot))
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%8.1f us %5.1f%% %4d  %s" % (v[1], 100.0 * v[1] / tot, v[0], k))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], tuple(sys.argv[3:]))   # further arguments: kernel-name substrings to leave out (the L2-flush fill between steps)

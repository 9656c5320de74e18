"""Per kernel AND grid size: average duration from an ncu launch list (`--metrics gpu__time_duration.sum --csv`).
Separates the multigrid levels, which run the same kernels on very different sizes.
    python profiles/by_grid.py gpurun_out/launches.csv [min_launches]"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
min_n = int(sys.argv[2]) if len(sys.argv) > 2 else 20
for i, r in enumerate(rows):
    if "Kernel Name" in r:
        hdr, start = r, i + 1
        break
ki, gi, vi, ui = hdr.index("Kernel Name"), hdr.index("Grid Size"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = collections.defaultdict(list)
for r in rows[start:]:
    if len(r) <= vi:
        continue
    try:
        v = float(r[vi].replace(",", ""))
    except ValueError:
        continue
    scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(r[ui], 1e-3)
    name = r[ki].split("(")[0].replace("mof::<unnamed>::", "").replace("void ", "")
    agg[(name, r[gi])].append(v * scale)
total = sum(sum(v) for v in agg.values())
print(f"# total {total / 1e3:.2f} ms; per (kernel, grid): launches, average us, total ms, share")
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    if len(v) < min_n:
        continue
    print(f"{k[0][:44]:44s} grid {k[1]:>14s} n={len(v):4d} avg={sum(v) / len(v):8.2f} us total={sum(v) / 1e3:7.2f} ms {100 * sum(v) / total:5.1f}%")

"""Turns gpurun_out/ ncu artefacts into the small text summaries committed under profiles/.
    python profiles/summarize.py launches gpurun_out/launches_TAG.csv > profiles/TAG_launches.txt
    python profiles/summarize.py kernel   gpurun_out/NAME_TAG.ncu-rep > profiles/TAG_NAME_ncu.txt
"""
import collections
import csv
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "l1tex__t_sector_hit_rate.pct",
        "lts__t_sector_hit_rate.pct", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "smsp__thread_inst_executed_per_inst_executed.ratio"]


def launches(path):
    lines = [ln for ln in open(path) if not ln.startswith("==")]
    tot = collections.defaultdict(lambda: [0, 0.0])
    for row in csv.DictReader(lines):
        v = float(row["Metric Value"].replace(",", ""))
        v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(row["Metric Unit"], 1.0)
        name = row["Kernel Name"].split("(")[0]
        tot[name][0] += 1
        tot[name][1] += v
    total = sum(v[1] for v in tot.values())
    print(f"# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: compare SHARES); total {total:.1f} ms over {sum(v[0] for v in tot.values())} launches")
    print(f"{'ms':>12s} {'share':>7s} {'launches':>9s}  kernel")
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        print(f"{v[1]:12.3f} {100 * v[1] / total:6.2f}% {v[0]:9d}  {k}")


def kernel(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    print(f"# ncu --set full --clock-control none, {path}")
    for row in data:
        print("kernel:", row[hdr.index("Kernel Name")][:100])
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                print(f"  {w:78s} {row[i]:>16s} {units[i]}")


if __name__ == "__main__":
    {"launches": launches, "kernel": kernel}[sys.argv[1]](sys.argv[2])

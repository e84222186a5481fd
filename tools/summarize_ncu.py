"""Writes the judged summaries under profiles/ from gpurun_out/ artefacts:
    python tools/summarize_ncu.py <tag> <report.ncu-rep>... [--launches launches.csv]"""
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__warps_eligible.avg.per_cycle_active", "sm__cycles_elapsed.max",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed"]


def summarize_report(path, out):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        out.write(f"## {r[hdr.index('Kernel Name')]}  grid {r[hdr.index('Grid Size')]} block {r[hdr.index('Block Size')]}\n")
        for k in KEYS:
            if k in hdr:
                out.write(f"  {k:62s} {r[hdr.index(k)]} {units[hdr.index(k)]}\n")
        stalls = {h.replace("smsp__pcsamp_warps_issue_stalled_", ""): float(r[i]) for i, h in enumerate(hdr)
                  if h.startswith("smsp__pcsamp_warps_issue_stalled_") and not h.endswith("_not_issued") and r[i]}
        tot = sum(stalls.values()) or 1.0
        out.write("  stall samples: " + ", ".join(f"{k} {v / tot:.0%}" for k, v in
                                                   sorted(stalls.items(), key=lambda x: -x[1])[:8]) + "\n\n")


def summarize_launches(path, out):
    agg = {}
    for row in csv.reader(open(path)):
        if len(row) > 14 and row[12] == "gpu__time_duration.sum":
            name = row[4].split("(")[0].replace("void ", "")
            a = agg.setdefault(name, [0, 0.0])
            a[0] += 1
            a[1] += float(row[14])
    total = sum(v[1] for v in agg.values()) or 1.0
    out.write(f"# launch list {path}: per-kernel count, mean duration, share of listed GPU time\n")
    for name, (n, ns) in sorted(agg.items(), key=lambda x: -x[1][1]):
        out.write(f"  {n:4d} x {ns / n / 1000:10.1f} us  {ns / total:6.1%}  {name[:110]}\n")
    out.write("\n")


if __name__ == "__main__":
    tag = sys.argv[1]
    args = sys.argv[2:]
    with open(f"profiles/{tag}.txt", "w") as out:
        out.write(f"# ncu summary {tag} (ncu --set full --clock-control none; cold-cache, serialised launches)\n\n")
        while args:
            a = args.pop(0)
            if a == "--launches":
                summarize_launches(args.pop(0), out)
            else:
                summarize_report(a, out)
    print(open(f"profiles/{tag}.txt").read())


def traffic_of(path, kernel_substr):
    """dram__bytes_read.sum + dram__bytes_write.sum (bytes) of the first kernel whose name contains
    `kernel_substr` in an `ncu --set full` report."""
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    for r in rows[2:]:
        if kernel_substr in r[hdr.index("Kernel Name")]:
            tot = 0.0
            for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                i = hdr.index(k)
                tot += float(r[i]) * scale[units[i]]
            return int(tot)
    return None

"""Stall samples of an `ncu --set full --import-source on` report, aggregated over the fill loops of the
file-resident sweep kernel (the instructions between its named-barrier sites) and over the whole kernel:

    python tools/ncu_regions.py report.ncu-rep > profiles/<name>.txt
"""
import csv
import subprocess
import sys


def main():
    path = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    name = rows[0][1]
    hdr, data = rows[1], rows[2:]
    ia, isrc = hdr.index("Address"), hdr.index("Source")
    ismp, iex = hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Instructions Executed")
    stalls = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    base = int(data[0][ia], 16)

    def region(lo, hi, title):
        sel = [r for r in data if lo <= int(r[ia], 16) - base < hi]
        tot = sum(int(r[ismp]) for r in sel)
        ex = sum(int(r[iex]) for r in sel)
        agg = sorted(((sum(int(r[i]) for r in sel), h[6:]) for i, h in stalls), reverse=True)
        if tot == 0:
            return
        print(f"{title}: {tot} samples, {ex} warp instructions")
        print("   " + ", ".join(f"{h} {n / tot:.0%}" for n, h in agg[:7] if n))

    print(f"# {path}\n# {name}")
    region(0, 1 << 40, "whole kernel (idle warps wait at the CTA barriers between phases)")
    bars = [int(r[ia], 16) - base for r in data if "BAR.SYNC" in r[isrc] and "0x1" in r[isrc]]
    clusters, cur = [], [bars[0]]
    for b in bars[1:]:
        if b - cur[-1] > 0x2000:
            clusters.append(cur)
            cur = [b]
        else:
            cur.append(b)
    clusters.append(cur)
    for c in clusters:
        region(c[0] - 0x400, c[-1] + 0x400, f"table fill instance at {hex(c[0])} ({len(c)} per-frame barrier sites)")


if __name__ == "__main__":
    main()

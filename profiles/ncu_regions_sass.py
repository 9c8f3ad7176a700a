"""Per-window instruction budget of the grouping kernel from an ncu capture: executed-per-window rate of every
SASS instruction (`ncu --page source --print-source sass`), contiguous runs with the same rate folded together.
usage: python profiles/ncu_regions_sass.py report.ncu-rep alignments_per_launch [alignments_per_window=30]"""
import csv
import subprocess
import sys


def main():
    rep, n_aln = sys.argv[1], float(sys.argv[2])
    per_window = float(sys.argv[3]) if len(sys.argv) > 3 else 30.0
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = next(r for r in rows if r and r[0] == "Address")
    data = rows[rows.index(hdr) + 1:]
    ia, isrc, ismp = hdr.index("Instructions Executed"), hdr.index("Source"), hdr.index("# Samples")
    windows = n_aln / per_window
    total = sum(int(r[ia]) for r in data)
    print("warp instructions %d, windows %.0f, per window %.1f" % (total, windows, total / windows))
    runs = []
    for i, r in enumerate(data):
        rate = int(r[ia]) / windows
        if runs and abs(runs[-1][2] - rate) < 0.06:
            runs[-1][1] = i
            runs[-1][3] += rate
            runs[-1][4] += int(r[ismp])
        else:
            runs.append([i, i, rate, rate, int(r[ismp]), r[isrc].strip()])
    print("%-13s %8s %6s %14s %8s  first instruction" % ("instructions", "rate", "count", "instr/window", "samples"))
    for a, b, rate, tot, smp, src in runs:
        if tot > 1.5:
            print("%5d-%-7d %8.2f %6d %14.1f %8d  %s" % (a, b, rate, b - a + 1, tot, smp, src[:60]))


if __name__ == "__main__":
    main()

"""Per-SASS-instruction view of one kernel in an .ncu-rep: executed warp instructions, stall samples.
usage: python profiles/ncu_regions.py report.ncu-rep [min_exec]  -> prints hot instructions and totals"""
import csv
import subprocess
import sys


def load(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = rows[1]
    ia, isrc, ismp, it = (hdr.index("Instructions Executed"), hdr.index("Source"), hdr.index("# Samples"),
                          hdr.index("Avg. Threads Executed"))
    res = []
    for r in rows[2:]:
        if len(r) < len(hdr) or not r[0].startswith("0x"):
            continue
        res.append((int(r[0], 16), int(r[ia]), int(r[ismp]), r[it], r[isrc].strip()))
    return res


def main():
    ins = load(sys.argv[1])
    min_exec = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    base = ins[0][0]
    tot_i = sum(i[1] for i in ins)
    tot_s = sum(i[2] for i in ins)
    print("warp instructions %d, stall samples %d" % (tot_i, tot_s))
    for a, n, s, t, src in ins:
        if n >= min_exec:
            print("%05x %10d %6d %5s  %s" % (a - base, n, s, t, src))


if __name__ == "__main__":
    main()

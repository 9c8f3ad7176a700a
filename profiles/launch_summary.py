"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel.
usage: python profiles/launch_summary.py launches.csv [launches_per_kernel_divisor]"""
import collections
import csv
import sys


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    hdr, agg, tot = None, collections.OrderedDict(), 0.0
    for r in rows:
        if hdr is None:
            if r and r[0] == "ID":
                hdr = r
            continue
        d = dict(zip(hdr, r))
        if d.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = d["Kernel Name"].split("(")[0].replace("void ", "")
        v = float(d["Metric Value"])
        v = v / 1e6 if d["Metric Unit"] == "ns" else (v / 1e3 if d["Metric Unit"] == "us" else v)
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
        tot += v
    print("%10s %6s %8s %7s  kernel" % ("total ms", "calls", "ms/call", "share"))
    for k, (c, v) in sorted(agg.items(), key=lambda x: -x[1][1]):
        print("%10.3f %6d %8.4f %6.1f%%  %s" % (v, c, v / c, 100 * v / tot, k))
    print("%10.3f ms in %d launches" % (tot, sum(c for c, _ in agg.values())))


if __name__ == "__main__":
    main()

"""Summarise `ncu --page source --csv --print-source cuda,sass` output per CUDA source line.
usage: python profiles/ncu_source_summary.py report.ncu-rep [top_n]"""
import csv
import subprocess
import sys


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 45
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    cur_file, cur_fn, hdr = None, None, None
    data = {}
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            cur_file = r[1].split("/")[-1]
            continue
        if r[0] == "Function Name":
            cur_fn = r[1]
            continue
        if r[0] == "Line No":
            hdr = r
            continue
        if hdr is None or len(r) < len(hdr) or r[2] != "-":
            continue  # keep only the per-line aggregate rows (Address == '-')
        d = dict(zip(hdr[4:], r[4:]))
        key = (cur_fn, cur_file, r[0], r[1].strip())
        inst = int(d["Instructions Executed"])
        smp = int(d["# Samples"])
        tinst = int(d["Thread Instructions Executed"])
        a = data.setdefault(key, [0, 0, 0])
        a[0] += inst
        a[1] += smp
        a[2] += tinst
    fns = sorted(set(k[0] for k in data))
    for fn in fns[:1]:
        items = [(k, v) for k, v in data.items() if k[0] == fn]
        tot = sum(v[0] for _, v in items)
        tots = sum(v[1] for _, v in items)
        print("kernel:", fn)
        print("warp instructions: %d   stall samples: %d" % (tot, tots))
        print("%7s %7s %6s  %s" % ("inst%", "smpl%", "thr/w", "file:line  source"))
        for k, v in sorted(items, key=lambda kv: -kv[1][0])[:top]:
            print("%6.2f%% %6.2f%% %6.1f  %s:%s  %s" % (100.0 * v[0] / tot, 100.0 * v[1] / max(tots, 1),
                                                       v[2] / max(v[0], 1), k[1], k[2], k[3][:95]))


if __name__ == "__main__":
    main()

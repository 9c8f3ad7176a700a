"""Warp instructions and stall samples of an ncu report summed over source-line ranges of one file.
usage: python profiles/ncu_line_ranges.py report.ncu-rep file.cuh name:lo-hi [name:lo-hi ...]"""
import csv, subprocess, sys
rep, fname = sys.argv[1], sys.argv[2]
ranges = []
for a in sys.argv[3:]:
    nm, r = a.split(":")
    lo, hi = r.split("-")
    ranges.append((nm, int(lo), int(hi)))
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur_file, hdr, first_fn, cur_fn = None, None, None, None
acc = {nm: [0, 0, 0] for nm, _, _ in ranges}
other = {}
tot = [0, 0, 0]
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]; continue
    if r[0] == "Function Name":
        cur_fn = r[1]; first_fn = first_fn or cur_fn; continue
    if r[0] == "Line No":
        hdr = r; continue
    if hdr is None or len(r) < len(hdr) or r[2] != "-" or cur_fn != first_fn:
        continue
    d = dict(zip(hdr[4:], r[4:]))
    v = (int(d["Instructions Executed"]), int(d["# Samples"]), int(d["Thread Instructions Executed"]))
    for i in range(3): tot[i] += v[i]
    line = int(r[0]); hit = False
    if cur_file == fname:
        for nm, lo, hi in ranges:
            if lo <= line <= hi:
                for i in range(3): acc[nm][i] += v[i]
                hit = True; break
    if not hit:
        o = other.setdefault(cur_file, [0, 0, 0])
        for i in range(3): o[i] += v[i]
print("total warp instr %d samples %d" % (tot[0], tot[1]))
for nm, v in list(acc.items()) + sorted(other.items(), key=lambda kv: -kv[1][0]):
    print("%-28s inst %6.2f%%  samples %6.2f%%  thr/w %5.1f" % (nm, 100.0 * v[0] / tot[0], 100.0 * v[1] / max(tot[1], 1), v[2] / max(v[0], 1)))

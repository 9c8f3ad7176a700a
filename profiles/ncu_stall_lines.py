"""Source lines of an ncu report ranked by one stall reason (default stall_long_sb).
usage: python profiles/ncu_stall_lines.py report.ncu-rep [stall_column] [top_n]"""
import csv, subprocess, sys
rep = sys.argv[1]
col = sys.argv[2] if len(sys.argv) > 2 else "stall_long_sb"
top = int(sys.argv[3]) if len(sys.argv) > 3 else 20
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = None; cur_file = None; data = {}; fn = None; first = None
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur_file = r[1].split("/")[-1]; continue
    if r[0] == "Function Name": fn = r[1]; first = first or fn; continue
    if r[0] == "Line No": hdr = r; continue
    if hdr is None or len(r) < len(hdr) or r[2] != "-" or fn != first: continue
    d = dict(zip(hdr[4:], r[4:]))
    key = (cur_file, int(r[0]), r[1].strip()[:90])
    a = data.setdefault(key, [0, 0]); a[0] += int(d[col] or 0); a[1] += int(d["# Samples"])
tot = sum(v[0] for v in data.values()); alls = sum(v[1] for v in data.values())
print("%s: %d of %d samples (%.1f%%)" % (col, tot, alls, 100.0 * tot / max(alls, 1)))
for k, v in sorted(data.items(), key=lambda kv: -kv[1][0])[:top]:
    print("%5.2f%%  %s:%d  %s" % (100.0 * v[0] / max(tot, 1), k[0], k[1], k[2]))

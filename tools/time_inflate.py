"""Pure block-decode speed of libbamcols' whole-buffer inflate against zlib on the BGZF blocks of a BAM file.
usage: python tools/time_inflate.py file.bam [max_blocks]"""
import ctypes, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from alntools_b200 import bam_io, bamcols

path = sys.argv[1]
limit = int(sys.argv[2]) if len(sys.argv) > 2 else 400
blocks = []
with open(path, "rb") as fh:
    data = fh.read()
off = 0
while off < len(data) and len(blocks) < limit:
    xlen = int.from_bytes(data[off + 10:off + 12], "little")
    bsize = int.from_bytes(data[off + 16:off + 18], "little") + 1
    isize = int.from_bytes(data[off + bsize - 4:off + bsize], "little")
    blocks.append((data[off + 12 + xlen:off + bsize - 8], isize))
    off += bsize
libs = [("built", bamcols.load_library())] + [(os.path.basename(p), bamcols.load_library(os.path.abspath(p))) for p in sys.argv[3:]]
total = sum(i for _, i in blocks)
buf = ctypes.create_string_buffer(1 << 17)
best = {}
for rep in range(7):
    for lname, lib in libs:
        for mode, name in ((1, "zlib"), (2, "fast")):
            t = time.perf_counter()
            for src, isize in blocks:
                rc = lib.bamcols_inflate_raw(src, len(src), buf, isize, mode)
                assert rc == 1 or isize == 0, rc
            dt = time.perf_counter() - t
            key = (lname, name)
            best[key] = min(best.get(key, 1e9), dt)
for (lname, name), dt in best.items():
    print("%-28s %s: %d blocks, %.1f MB inflated, best of 7: %.0f MB/s" % (lname, name, len(blocks), total / 1e6, total / 1e6 / dt))

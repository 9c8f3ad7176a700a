"""Timing of the per-cell host emitter (parallel window code vs BAMCOLS_SEQUENTIAL_CELLS) on a synthetic
10x-style BAM with fixed-length '|||' names.  Usage: python tools/time_cells.py [n_reads]"""
import os, sys, time, struct
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from alntools_b200 import bam_io, bamcols, synth
from alntools_b200.header import TargetTables

n_reads = int(sys.argv[1]) if len(sys.argv) > 1 else 2000000
T, H, CELLS = 2000, 2, 5000
cols = synth.make_columns(n_reads, T, H, seed=3, mode="light")
rg = cols["read_group"].astype(np.int64)
n = len(rg)
rng = np.random.default_rng(5)
cell_of_read = rng.integers(0, CELLS, size=int(rg.max()) + 1)
refs = synth.reference_names(T, H)
tmpl = ("r%08d" % 0) + "|||x" * 13 + "|||C%05d" % 0
name_len = len(tmpl) + 1
rec_len = 36 + name_len
raw = np.zeros((n, rec_len), dtype=np.uint8)
raw[:, 0:4] = np.frombuffer(struct.pack("<i", 32 + name_len), dtype=np.uint8)
core = np.zeros(n, dtype=np.dtype([("tid", "<i4"), ("pos", "<i4"), ("l_name", "u1"), ("mapq", "u1"), ("bin", "<u2"),
                                   ("ncig", "<u2"), ("flag", "<u2"), ("lseq", "<i4"), ("ntid", "<i4"), ("npos", "<i4"),
                                   ("tlen", "<i4")]))
core["tid"] = cols["target_idx"].astype(np.int64) * H + cols["hap_idx"]
core["l_name"] = name_len
core["ntid"] = -1
core["npos"] = -1
raw[:, 4:36] = core.view(np.uint8).reshape(n, 32)
raw[:, 36:36 + len(tmpl)] = np.frombuffer(tmpl.encode(), dtype=np.uint8)
rem = rg.copy()
for d in range(8, 0, -1):
    raw[:, 36 + d] = (rem % 10 + 48).astype(np.uint8); rem //= 10
rem = cell_of_read[rg].copy()
for d in range(len(tmpl) - 1, len(tmpl) - 6, -1):
    raw[:, 36 + d] = (rem % 10 + 48).astype(np.uint8); rem //= 10
flat = raw.reshape(-1)
per_block = max(1, 60000 // rec_len) * rec_len
path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "gpurun_out", "tmp", "cells.bam")
with open(path, "wb") as fh:
    bam_io._write_blocks(fh, bam_io.bam_header_bytes(refs), 60000, 1)
    for off in range(0, flat.size, per_block):
        fh.write(bam_io.bgzf_block(flat[off:off + per_block].tobytes(), 1))
    fh.write(bam_io.BGZF_EOF)
tables = TargetTables([r[0] for r in refs], [r[1] for r in refs], None)
res = {}
for mode in ("parallel", "sequential", "parallel"):
    if mode == "sequential":
        os.environ["BAMCOLS_SEQUENTIAL_CELLS"] = "1"
    else:
        os.environ.pop("BAMCOLS_SEQUENTIAL_CELLS", None)
    cells = bamcols.CellDictionary()
    t0 = time.time()
    with bamcols.BamColumnReader(path) as r:
        r.set_tables(tables)
        got = r.read_all(cells=cells, chunk=1 << 22)
        ph = r.phase_seconds()
    dt = time.time() - t0
    print("%-10s %.3f s  %.1f M alignments/s  (%d rows, %d cells)" % (mode, dt, n / dt / 1e6, len(got["read_group"]), len(cells.names())))
    print("   ", " ".join("%s=%.3f" % kv for kv in ph.items()))
    res[mode] = got
for k in res["parallel"]:
    assert np.array_equal(res["parallel"][k], res["sequential"][k]), k
print("identical columns")

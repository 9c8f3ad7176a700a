"""Mint the --rangefile goldens (tests/golden/range_*) by running the UNMODIFIED reference.

TEST INFRASTRUCTURE ONLY, build-container only.  Kept apart from make_golden.py so that the EC-file
goldens do not have to be re-minted.  The range file (alntools/bam_utils.py:282-286,735-766;
bam_utils_multisample.py:249-253,568-576,638-668) holds, per main target and haplotype,
max(reference_start) - min(reference_start) + 1 over the valid alignments, or 0."""
import glob
import json
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from alntools_b200 import bam_io  # noqa: E402
from oracle import run_reference  # noqa: E402
from oracle.make_golden import cid_name, sha256  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def main():
    if not run_reference.available():
        sys.exit("the reference is not present; goldens can only be minted in the build container")
    rng = np.random.default_rng(17)
    refs = [("T0_A", 900), ("T0_B", 901), ("T1_A", 500), ("T2", 700), ("T3_A", 400), ("T3_B", 401), ("T4_B", 300)]
    alns = []
    for read in range(120):
        name = "r%03d" % read if read % 7 else "r%03d with blank" % read
        for _ in range(int(rng.integers(1, 5))):
            flag = int(rng.choice([0, 0, 0, 16, 4, 1 | 2 | 64, 1 | 2 | 128, 1 | 64]))
            tid = int(rng.choice([0, 1, 2, 3, 4, 5]))           # T4_B never gets an alignment
            pos = int(rng.integers(0, 800))
            alns.append((name, flag, tid, pos, tid, int(rng.choice([-1, 40]))))
    manifest = []
    tmp = tempfile.mkdtemp(prefix="golden_range_")
    bam = os.path.join(GOLD, "range_single.bam")
    bam_io.write_bam(bam, refs, alns, block_payload=900)
    targets = os.path.join(GOLD, "range_single.targets.txt")
    with open(targets, "w") as fh:
        fh.write("# targets first\nT3\nT9 extra\n")
    for tag, tfile in (("range_single", None), ("range_single_targets", targets)):
        out, rfile = os.path.join(GOLD, tag + ".bin"), os.path.join(GOLD, tag + ".range.txt")
        run_reference.bam2ec(bam, out, 1, 1, tfile, temp_dir=tmp, range_filename=rfile)
        manifest.append({"name": tag, "kind": "single", "bam": "range_single.bam",
                         "targets": os.path.basename(tfile) if tfile else None, "ec": tag + ".bin",
                         "range": tag + ".range.txt", "sha256": sha256(out), "reference_chunks_checked": [1]})

    d = os.path.join(GOLD, "range_multi")
    os.makedirs(d, exist_ok=True)
    for f in range(2):
        falns = []
        for read in range(60):
            cell = "CELL%d" % int(rng.integers(0, 4))
            for _ in range(int(rng.integers(1, 4))):
                falns.append((cid_name("f%dq%03d" % (f, read), cell), 0, int(rng.integers(0, 6)), int(rng.integers(0, 300))))
        bam_io.write_bam(os.path.join(d, "part%d.bam" % f), refs, falns, block_payload=900)
    order = [os.path.basename(p) for p in glob.glob(os.path.join(d, "*.bam"))]
    out, rfile = os.path.join(GOLD, "range_multi.m1.bin"), os.path.join(GOLD, "range_multi.range.txt")
    run_reference.bam2ec_multisample(d, out, 1, range_filename=rfile)
    manifest.append({"name": "range_multi", "kind": "multisample", "dir": "range_multi", "file_order": order,
                     "mincount": 1, "ec": "range_multi.m1.bin", "range": "range_multi.range.txt",
                     "sha256": sha256(out)})
    with open(os.path.join(GOLD, "manifest_range.json"), "w") as fh:
        json.dump(manifest, fh, indent=1)
    print("wrote", [m["name"] for m in manifest])


if __name__ == "__main__":
    main()

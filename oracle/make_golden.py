"""Mint tests/golden/: synthetic BAMs + the EC files the UNMODIFIED reference writes for them.

TEST INFRASTRUCTURE ONLY.  Run in the build container (needs /root/reference):

    python oracle/make_golden.py

Every case is listed in tests/golden/manifest.json (inputs, arguments, expected output, sha256).
Multi-chunk runs of the reference (with the PEP-479 harness patch) are checked to be byte-identical
to the 1-chunk run before the golden is accepted.
"""
import hashlib
import json
import os
import shutil
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from alntools_b200 import bam_io, synth  # noqa: E402
from oracle import run_reference  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def cid_name(orig, cell):
    """10x-style read name whose '|||' field 14 is the cell id (bam_utils_multisample.py:270-280)."""
    return "|||".join([orig, "CR", "x", "CY", "x", "UR", "x", "UY", "x", "BC", "x", "QT", "x", "CID", cell])


def case_toy_single():
    refs = [("T0_A", 100), ("T0_B", 101), ("T1_A", 200), ("T1_B", 201), ("T10_A", 300), ("T10_B", 301)]
    alns = [("r1", 0, 0), ("r1", 0, 1), ("r2", 0, 2), ("r3", 0, 2), ("r3", 0, 5),
            ("r4", 0, 1), ("r4", 0, 0), ("r5", 0, 2), ("r6", 4, -1), ("r7", 0, 0), ("r7", 0, 0), ("r7", 0, 1)]
    return refs, alns


def case_mixed_single():
    """'' haplotype next to A/B, a leading-underscore name, names with spaces, paired-end filters,
    unmapped records, duplicate tids, target file that reorders targets."""
    refs = [("G1_A", 10), ("G1_B", 11), ("G2", 20), ("_G3", 30), ("G4_B", 41), ("G2_A", 21), ("G5_x_A", 50)]
    P, PP, R2, UN = 0x1, 0x2, 0x80, 0x4
    alns = [
        ("q1 extra/1", 0, 0), ("q1 other", 0, 1), ("q1", 0, 2),
        ("q2", P | PP, 3, 5, 3, 40), ("q2", P | PP | R2, 3, 40, 3, 5),
        ("q3", P, 4, 5, 4, 40),                      # not proper pair -> dropped (whole read vanishes)
        ("q4", P | PP, 5, 5, 2, 40),                 # mate on another reference -> dropped
        ("q4", P | PP, 5, 5, 5, -1),                 # mate unplaced -> dropped
        ("q4", P | PP, 6, 7, 6, 70),
        ("q5", UN, -1),
        ("q5", 0, 2), ("q5", 0, 5), ("q5", 0, 2),
        (" q6", 0, 0),                               # space at index 0 is NOT trimmed
        (" q6", 0, 1),
        ("q7", 0, 2), ("q7", 0, 5),                  # same EC as q5 via a different order
        ("q8", 0, 6), ("q9", 0, 3),
    ]
    targets = "# comment line\nG9 ignored-second-token\nG2\nG1\n"
    return refs, alns, targets


def case_multisample():
    refs = [("T0_A", 100), ("T0_B", 101), ("T1_A", 200), ("T1_B", 201), ("T2_A", 300), ("T2_B", 301)]
    a = [(cid_name("a1", "cellB"), 0, 0), (cid_name("a1", "cellB"), 0, 1),
         (cid_name("a2", "cellA"), 0, 2),
         (cid_name("a3", "cellB"), 0, 2),
         (cid_name("a4", "cellC"), 0, 0), (cid_name("a4", "cellC"), 0, 1), (cid_name("a4", "cellC"), 0, 0),
         (cid_name("a5", "cellA"), 0, 4),
         (cid_name("a6", "cellA"), 4, -1),
         (cid_name("a7", "cellD"), 0, 5),            # last read of the file: dropped
         ]
    b = [(cid_name("b1", "cellD"), 0, 2),
         (cid_name("b2", "cellA"), 0, 1), (cid_name("b2", "cellA"), 0, 0),
         (cid_name("b3", "cellE sp"), 0, 3), (cid_name("b3", "cellE sp"), 0, 2),  # space quirk
         (cid_name("b4", "cellD"), 0, 2),
         (cid_name("b5", "cellB"), 0, 4),
         (cid_name("b6", "cellB"), 0, 4),
         ]
    return refs, [("a.bam", a), ("b.bam", b)]


def synthetic_multisample_files(seed, n_files, reads_per_file, n_targets, n_haps, n_cells):
    files = []
    refs = synth.reference_names(n_targets, n_haps)
    for f in range(n_files):
        cols = synth.make_columns(reads_per_file, n_targets, n_haps, seed * 100 + f, mode="light",
                                  n_cells=n_cells, dup_rate=0.02)
        tids = cols["target_idx"].astype(np.int64) * n_haps + cols["hap_idx"]
        alns = [(cid_name("f%dr%07d" % (f, rg), "cell%04d" % c), 0, int(t))
                for rg, c, t in zip(cols["read_group"].tolist(), cols["cell_idx"].tolist(), tids.tolist())]
        files.append(("s%d.bam" % f, alns))
    return refs, files


def sha256(path):
    with open(path, "rb") as fh:
        return hashlib.sha256(fh.read()).hexdigest()


def main():
    if not run_reference.available():
        sys.exit("the reference is not present; goldens can only be minted in the build container")
    if os.path.isdir(GOLD):
        shutil.rmtree(GOLD)
    os.makedirs(GOLD)
    manifest = []
    tmp = tempfile.mkdtemp(prefix="golden_")

    def single(name, refs, alns, targets=None, check_chunks=()):
        bam = os.path.join(GOLD, name + ".bam")
        bam_io.write_bam(bam, refs, alns, block_payload=4000)
        tfile = None
        if targets is not None:
            tfile = os.path.join(GOLD, name + ".targets.txt")
            with open(tfile, "w") as fh:
                fh.write(targets)
        out = os.path.join(GOLD, name + ".bin")
        run_reference.bam2ec(bam, out, 1, 1, tfile, temp_dir=tmp)
        for nc in check_chunks:
            alt = os.path.join(tmp, "alt.bin")
            run_reference.bam2ec(bam, alt, nc, min(nc, 4), tfile, temp_dir=tmp)
            assert sha256(alt) == sha256(out), "reference output depends on chunk count (%d)" % nc
        manifest.append({"name": name, "kind": "single", "bam": name + ".bam",
                         "targets": os.path.basename(tfile) if tfile else None,
                         "ec": name + ".bin", "sha256": sha256(out),
                         "reference_chunks_checked": [1] + list(check_chunks)})

    def multi(name, refs, files, mincounts):
        import glob
        d = os.path.join(GOLD, name)
        os.makedirs(d)
        for fn, alns in files:
            bam_io.write_bam(os.path.join(d, fn), refs, alns, block_payload=4000)
        for mincount in mincounts:
            out = os.path.join(GOLD, "%s.m%d.bin" % (name, mincount))
            # the reference merges files in glob order (bam_utils_multisample.py:379): record it
            order = [os.path.basename(p) for p in glob.glob(os.path.join(d, "*.bam"))]
            run_reference.bam2ec_multisample(d, out, mincount)
            manifest.append({"name": "%s.m%d" % (name, mincount), "kind": "multisample", "dir": name,
                             "file_order": order, "mincount": mincount,
                             "ec": os.path.basename(out), "sha256": sha256(out)})

    refs, alns = case_toy_single()
    single("toy_single", refs, alns)
    refs, alns, targets = case_mixed_single()
    single("mixed_single", refs, alns)
    single("mixed_single_targets", refs, alns, targets)

    cols = synth.make_columns(20000, 500, 2, seed=1, mode="light", dup_rate=0.02)
    tids = cols["target_idx"].astype(np.int64) * 2 + cols["hap_idx"]
    alns = [("read%09d" % rg, 0, int(t)) for rg, t in zip(cols["read_group"].tolist(), tids.tolist())]
    single("synth_h2", synth.reference_names(500, 2), alns, check_chunks=(3, 8))

    cols = synth.make_columns(3000, 300, 8, seed=3, mode="heavy", dup_rate=0.01)
    tids = cols["target_idx"].astype(np.int64) * 8 + cols["hap_idx"]
    alns = [("read%09d" % rg, 0, int(t)) for rg, t in zip(cols["read_group"].tolist(), tids.tolist())]
    single("synth_h8_heavy", synth.reference_names(300, 8), alns, check_chunks=(4,))

    refs, files = case_multisample()
    multi("toy_multi", refs, files, (-1, 2, 3))

    refs, files = synthetic_multisample_files(4, 3, 4000, 200, 2, 40)
    multi("synth_multi", refs, files, (1, 150))

    with open(os.path.join(GOLD, "manifest.json"), "w") as fh:
        json.dump(manifest, fh, indent=1)
    shutil.rmtree(tmp)
    total = sum(os.path.getsize(os.path.join(dp, f)) for dp, _, fs in os.walk(GOLD) for f in fs)
    print("wrote %d cases, %d bytes under %s" % (len(manifest), total, GOLD))


if __name__ == "__main__":
    main()

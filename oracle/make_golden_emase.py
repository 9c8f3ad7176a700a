"""Mint tests/golden/*.emase.pkl: everything the UNMODIFIED reference hands to PyTables when bam2emase /
bam2ec writes its EMASE (.h5) file (Sparse3DMatrix.save + AlignmentPropertyMatrix.save), recorded by the
`tables` shim (oracle/shims/tables.py).  TEST INFRASTRUCTURE ONLY, build container only.

PyTables itself is not installable here, so the bytes of an .h5 file cannot be compared; the sequence of
groups, arrays (dtype, shape, values), titles, attributes and filter settings can.
usage: python oracle/make_golden_emase.py
"""
import json
import os
import pickle
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLDEN = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, ROOT)

from oracle import run_reference   # noqa: E402


def record_reference(case):
    bam_utils, multi = run_reference._import_reference()
    run_reference.patch_pep479(bam_utils)
    import tables                        # the shim
    rec = tables.start_recording()
    tmp = tempfile.mkdtemp()
    h5 = os.path.join(tmp, "out.h5")
    try:
        if case["kind"] == "single":
            tfile = os.path.join(GOLDEN, case["targets"]) if case["targets"] else None
            # the reference names the sample after the file: keep the golden's own file name
            bam_utils.convert(os.path.join(GOLDEN, case["bam"]), None, h5, num_chunks=1, number_processes=1,
                              temp_dir=tmp, target_filename=tfile)
        else:
            # the reference globs the directory; the goldens were minted with that order (file_order)
            multi.convert(os.path.join(GOLDEN, case["dir"]), None, h5, 0, case["mincount"], 1, tmp, None, None)
    finally:
        tables.stop_recording()
    return rec[h5]


def main():
    with open(os.path.join(GOLDEN, "manifest.json")) as fh:
        cases = json.load(fh)
    for case in cases:
        events = record_reference(case)
        out = os.path.join(GOLDEN, case["name"] + ".emase.pkl")
        with open(out, "wb") as fh:
            pickle.dump(events, fh, protocol=2)
        print("%-28s %3d events, %d arrays -> %s" % (case["name"], len(events), sum(e["op"] == "carray" for e in events),
                                                    os.path.relpath(out, ROOT)))


if __name__ == "__main__":
    main()

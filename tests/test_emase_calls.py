"""EMASE (.h5) output pinned at the PyTables call level.  PyTables cannot be
installed in this image, so no .h5 bytes exist to compare; what the writers hand to PyTables can be: the
unmodified reference's APM.save() was run against a recording `tables` module (oracle/shims/tables.py,
oracle/make_golden_emase.py -> tests/golden/*.emase.pkl) and alntools_b200.emase.save_emase(), fed from the
golden EC file of the same case, must leave the same record: groups, arrays (dtype, shape, values), titles,
attributes and filter settings, in the same order."""
import importlib.util
import os
import pickle
import sys

import numpy as np
import pytest

from conftest import GOLDEN, golden_cases
from alntools_b200 import bin_utils, emase

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _recording_tables():
    spec = importlib.util.spec_from_file_location("tables", os.path.join(ROOT, "oracle", "shims", "tables.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _same_event(got, want, where):
    assert got["op"] == want["op"], where
    for key in want:
        if key == "array":
            g, w = np.asarray(got["array"]), np.asarray(want["array"])
            assert g.dtype == w.dtype and g.shape == w.shape, (where, g.dtype, w.dtype, g.shape, w.shape)
            assert np.array_equal(g, w), where
        else:
            assert got[key] == want[key], (where, key, got[key], want[key])


CASES = [c for c in golden_cases("single") + golden_cases("multisample")
         if os.path.isfile(os.path.join(GOLDEN, c["name"] + ".emase.pkl"))]


def test_every_ec_golden_has_its_emase_record():
    import json
    with open(os.path.join(GOLDEN, "manifest.json")) as fh:
        names = [c["name"] for c in json.load(fh)]
    assert sorted(c["name"] for c in CASES) == sorted(names) and len(names) == 10


@pytest.mark.parametrize("case", CASES, ids=lambda c: c["name"])
def test_save_emase_hands_pytables_what_the_reference_does(case, tmp_path, monkeypatch):
    with open(os.path.join(GOLDEN, case["name"] + ".emase.pkl"), "rb") as fh:
        want = pickle.load(fh)
    ec = bin_utils.ecload_arrays(os.path.join(GOLDEN, case["ec"]))
    tables = _recording_tables()
    monkeypatch.setitem(sys.modules, "tables", tables)
    rec = tables.start_recording()
    h5 = str(tmp_path / "out.h5")
    single = case["kind"] == "single"
    shape = (len(ec["targets"]), len(ec["haplotypes"]), len(ec["a"][0]) - 1)
    try:
        emase.save_emase(h5, "bam2ec" if single else "Multisample APM", shape, ec["haplotypes"], ec["targets"],
                         ec["lengths"], ec["samples"], ec["a"], ec["n"], incidence_only=single)
    finally:
        tables.stop_recording()
    got = rec[h5]
    # the reference writes in two sessions (Sparse3DMatrix.save, then APM.save re-opens in append mode);
    # session boundaries are not part of the file: compare what is created, in order
    strip = lambda events: [e for e in events if e["op"] not in ("open", "close")]
    titles = lambda events: [e["title"] for e in events if e["op"] == "open" and e["mode"] == "w"]
    assert titles(got) == titles(want)
    got, want = strip(got), strip(want)
    assert [(e["op"], e.get("node"), e.get("name")) for e in got] == [(e["op"], e.get("node"), e.get("name")) for e in want]
    for i, (g, w) in enumerate(zip(got, want)):
        _same_event(g, w, "%s event %d %s %s" % (case["name"], i, w["op"], w.get("node")))


def test_reference_reproduces_the_committed_records():
    """Where the reference is present (the build container), running it again against the recording shim must
    give the committed records - they are what the reference does, not what this repository thinks it does."""
    sys.path.insert(0, ROOT)
    from oracle import run_reference
    if not run_reference.available():
        pytest.skip("the reference tree is not on this machine")
    from oracle import make_golden_emase
    import json
    with open(os.path.join(GOLDEN, "manifest.json")) as fh:
        cases = json.load(fh)
    for case in cases[:3] + cases[5:7]:          # three single-sample and two per-cell cases keep this quick
        with open(os.path.join(GOLDEN, case["name"] + ".emase.pkl"), "rb") as fh:
            want = pickle.load(fh)
        got = make_golden_emase.record_reference(case)
        assert len(got) == len(want), case["name"]
        for i, (g, w) in enumerate(zip(got, want)):
            _same_event(g, w, "%s event %d" % (case["name"], i))


def test_ec2emase_hands_pytables_what_the_reference_does(tmp_path, monkeypatch):
    """`ec2emase` (bin_utils.py:979-995): the reference reads the EC file back and saves it with the data arrays
    included; ours must leave the same PyTables record for the same file."""
    sys.path.insert(0, ROOT)
    from oracle import run_reference
    if not run_reference.available():
        pytest.skip("the reference tree is not on this machine")
    run_reference._import_reference()
    from alntools import bin_utils as ref_bin_utils          # the reference package
    import tables as shim                                     # the recording shim the reference imported
    strip = lambda events: [e for e in events if e["op"] not in ("open", "close")]
    for case in CASES[:2] + CASES[5:7]:
        ec_file = os.path.join(GOLDEN, case["ec"])
        rec = shim.start_recording()
        ref_h5 = str(tmp_path / (case["name"] + ".ref.h5"))
        try:
            ref_bin_utils.ec2emase(ec_file, ref_h5)
        finally:
            shim.stop_recording()
        want = strip(rec.get(ref_h5, []))
        assert want, "the reference wrote nothing for %s (its ecload failed?)" % case["name"]
        mine = _recording_tables()
        monkeypatch.setitem(sys.modules, "tables", mine)
        rec2 = mine.start_recording()
        our_h5 = str(tmp_path / (case["name"] + ".our.h5"))
        try:
            bin_utils.ec2emase(ec_file, our_h5)
        finally:
            mine.stop_recording()
        monkeypatch.setitem(sys.modules, "tables", shim)
        got = strip(rec2[our_h5])
        assert [(e["op"], e.get("node"), e.get("name")) for e in got] == [(e["op"], e.get("node"), e.get("name")) for e in want]
        for i, (g, w) in enumerate(zip(got, want)):
            _same_event(g, w, "%s event %d %s %s" % (case["name"], i, w["op"], w.get("node")))

"""Pin the oracle: it must reproduce, byte for byte, the EC files written by the UNMODIFIED reference
(tests/golden/, minted by oracle/make_golden.py) and agree with itself across its three forms
(record-level Python, column-level numpy, column-level C)."""
import hashlib
import os

import numpy as np
import pytest

from conftest import GOLDEN, golden_cases, load_records
from oracle import ec_oracle


def _sha(b):
    return hashlib.sha256(b).hexdigest()


@pytest.mark.parametrize("case", golden_cases("single"), ids=lambda c: c["name"])
def test_record_oracle_matches_reference_single(case):
    header, recs = load_records(os.path.join(GOLDEN, case["bam"]))
    tfile = os.path.join(GOLDEN, case["targets"]) if case["targets"] else None
    got = ec_oracle.convert_single(header.references, header.lengths, [recs], case["bam"], tfile)
    with open(os.path.join(GOLDEN, case["ec"]), "rb") as fh:
        want = fh.read()
    assert _sha(want) == case["sha256"]
    assert got == want


@pytest.mark.parametrize("case", golden_cases("single"), ids=lambda c: c["name"])
@pytest.mark.parametrize("n_chunks", [2, 5])
def test_record_oracle_chunk_independent(case, n_chunks):
    """bam_utils.py:680-698: merging chunk results in order gives the same file as one chunk, as long
    as reads do not span chunks."""
    header, recs = load_records(os.path.join(GOLDEN, case["bam"]))
    names = [ec_oracle.trim_name(r[0]) for r in recs]
    cuts = [0]
    for k in range(1, n_chunks):
        c = len(recs) * k // n_chunks
        while 0 < c < len(recs) and names[c] == names[c - 1]:
            c += 1
        cuts.append(max(c, cuts[-1]))
    cuts.append(len(recs))
    chunks = [recs[a:b] for a, b in zip(cuts[:-1], cuts[1:]) if b > a]
    chunks = [c for c in chunks if any(ec_oracle.alignment_is_valid(r[1], r[2], r[4], r[5]) for r in c)]
    tfile = os.path.join(GOLDEN, case["targets"]) if case["targets"] else None
    got = ec_oracle.convert_single(header.references, header.lengths, chunks, case["bam"], tfile)
    with open(os.path.join(GOLDEN, case["ec"]), "rb") as fh:
        assert got == fh.read()


@pytest.mark.parametrize("case", golden_cases("multisample"), ids=lambda c: c["name"])
def test_record_oracle_matches_reference_multisample(case):
    files = []
    header0 = None
    for fn in case["file_order"]:
        header, recs = load_records(os.path.join(GOLDEN, case["dir"], fn))
        header0 = header0 or header
        files.append(recs)
    got = ec_oracle.convert_multisample(header0.references, header0.lengths, files, case["mincount"])
    with open(os.path.join(GOLDEN, case["ec"]), "rb") as fh:
        want = fh.read()
    assert _sha(want) == case["sha256"]
    assert got == want


def test_worked_example_from_survey():
    """reads {T0_A,T0_B}x2, {T1_A}x2, {T1_A,T10_B}x1."""
    rg = [0, 0, 1, 2, 2, 3, 3, 4]
    tg = [0, 0, 1, 1, 2, 0, 0, 1]
    hp = [0, 1, 0, 0, 1, 1, 0, 0]
    indptr, indices, data, counts = ec_oracle.ec_from_columns(rg, tg, hp)
    assert indptr.tolist() == [0, 1, 2, 4]
    assert indices.tolist() == [0, 1, 1, 2]
    assert data.tolist() == [3, 1, 1, 2]
    assert counts.tolist() == [2, 2, 1]


@pytest.mark.parametrize("seed,mode,haps", [(11, "light", 2), (12, "diploid", 2), (13, "heavy", 8), (14, 64, 8)])
def test_c_oracle_matches_numpy_oracle(built, seed, mode, haps):
    from alntools_b200 import synth
    from oracle import c_oracle
    cols = synth.make_columns(3000 if mode != 64 else 300, 400, haps, seed, mode=mode, dup_rate=0.03)
    want = ec_oracle.ec_from_columns(cols["read_group"], cols["target_idx"], cols["hap_idx"])
    got = c_oracle.ec_from_columns(cols["read_group"], cols["target_idx"], cols["hap_idx"])
    for a, b in zip(got[:4], want):
        assert np.array_equal(a, b)
    assert got[4] == cols["n_reads"]
    assert int(got[3].sum()) == cols["n_reads"]


def test_c_oracle_edge_cases(built):
    from oracle import c_oracle
    e = np.zeros(0, dtype=np.int32)
    indptr, indices, data, counts, n_reads = c_oracle.ec_from_columns(e, e, e)
    assert indptr.tolist() == [0] and n_reads == 0 and len(counts) == 0
    one = np.array([5], dtype=np.int32)
    indptr, indices, data, counts, n_reads = c_oracle.ec_from_columns(one, one, np.array([3], dtype=np.int32))
    assert (indptr.tolist(), indices.tolist(), data.tolist(), counts.tolist()) == ([0, 1], [5], [8], [1])
    indptr, *_rest, n_reads = c_oracle.ec_from_columns(one, one, one, drop_last=True)
    assert indptr.tolist() == [0] and n_reads == 0


@pytest.mark.skipif(not os.path.isdir("/root/reference/alntools"), reason="reference only exists in the build container")
def test_goldens_regenerate_identically(tmp_path):
    """Re-run the unmodified reference on one golden input and compare with the committed file."""
    from oracle import run_reference
    case = golden_cases("single")[0]
    out = str(tmp_path / "out.bin")
    bam = str(tmp_path / case["bam"])
    with open(os.path.join(GOLDEN, case["bam"]), "rb") as src, open(bam, "wb") as dst:
        dst.write(src.read())
    run_reference.bam2ec(bam, out, 1, 1, None, temp_dir=str(tmp_path))
    with open(out, "rb") as a, open(os.path.join(GOLDEN, case["ec"]), "rb") as b:
        assert a.read() == b.read()


def test_reference_ecfile_reader_decodes_our_files_like_we_do(tmp_path):
    """the reference's own ECFile reader (bin_file.py:99-402) as an independent
    decoder.  Every golden EC file, and the same content written again by bin_utils.ecsave2_arrays, must
    come back from it with the names and matrices that bin_utils.ecload_arrays reads."""
    from oracle import run_reference
    if not run_reference.available():
        pytest.skip("the reference tree is not on this machine")
    run_reference._import_reference()
    from alntools import bin_file                      # the reference package
    from alntools_b200 import bin_utils
    import json
    with open(os.path.join(GOLDEN, "manifest.json")) as fh:
        cases = json.load(fh)
    for case in cases:
        src = os.path.join(GOLDEN, case["ec"])
        ours = bin_utils.ecload_arrays(src)
        again = str(tmp_path / (case["name"] + ".again.bin"))
        bin_utils.ecsave2_arrays(again, ours["haplotypes"], ours["targets"], ours["lengths"], ours["samples"],
                                 ours["a"], ours["n"])
        with open(src, "rb") as a, open(again, "rb") as b:
            assert a.read() == b.read(), case["name"]
        ref = bin_file.ECFile(again)
        text = lambda v: v.decode() if isinstance(v, bytes) else v   # names come back as bytes on py3
        assert [text(h) for h in ref.haplotypes_idx] == ours["haplotypes"], case["name"]
        assert [text(t) for t in ref.targets_idx] == ours["targets"], case["name"]
        assert [text(s) for s in ref.samples_idx] == ours["samples"], case["name"]
        assert np.array_equal(ref.a_matrix.indptr, ours["a"][0]) and np.array_equal(ref.a_matrix.indices, ours["a"][1])
        assert np.array_equal(ref.a_matrix.data, ours["a"][2]), case["name"]
        assert np.array_equal(ref.n_matrix.indptr, ours["n"][0]) and np.array_equal(ref.n_matrix.indices, ours["n"][1])
        assert np.array_equal(ref.n_matrix.data, ours["n"][2]), case["name"]
        assert ref.a_matrix.shape[0] == len(ours["a"][0]) - 1


def test_reference_writers_accept_our_output_object(tmp_path):
    """SURVEY section 8b, 'Output object': alntools_b200.apm.ApmArrays carries the fields the reference's writers
    read.  Built from the arrays of every golden EC file and handed to the UNMODIFIED reference's
    bin_utils.ecsave2, it must produce the golden bytes again; handed to the reference's APM.save code path
    (run against the recording `tables` shim) it must leave the committed EMASE record."""
    from oracle import run_reference
    if not run_reference.available():
        pytest.skip("the reference tree is not on this machine")
    run_reference._import_reference()
    from alntools import bin_utils as ref_bin_utils          # the reference package
    from alntools.matrix.AlignmentPropertyMatrix import AlignmentPropertyMatrix as RefAPM
    from alntools_b200 import bin_utils
    from alntools_b200.apm import ApmArrays
    import json
    import pickle
    import tables                                            # the shim (on sys.path after _import_reference)
    with open(os.path.join(GOLDEN, "manifest.json")) as fh:
        cases = json.load(fh)
    for case in cases:
        src = os.path.join(GOLDEN, case["ec"])
        ec = bin_utils.ecload_arrays(src)
        apm = ApmArrays(ec["haplotypes"], ec["targets"], ec["lengths"], ec["samples"], ec["a"], ec["n"])
        # the reference's save() as an unbound function over our object (convert() calls it before ecsave2,
        # bam_utils.py:849-876, and ecsave2 widens apm.lengths in place)
        single = case["kind"] == "single"
        rec = tables.start_recording()
        h5 = str(tmp_path / (case["name"] + ".h5"))
        try:
            RefAPM.save(apm, h5, title="bam2ec" if single else "Multisample APM", incidence_only=single)
        finally:
            tables.stop_recording()
        with open(os.path.join(GOLDEN, case["name"] + ".emase.pkl"), "rb") as fh:
            want = pickle.load(fh)
        got = rec[h5]
        assert len(got) == len(want), case["name"]
        for g, w in zip(got, want):
            assert g["op"] == w["op"] and g.get("node") == w.get("node"), case["name"]
            if g["op"] == "carray":
                assert str(np.asarray(g["array"]).dtype) == str(np.asarray(w["array"]).dtype), (case["name"], g["node"])
                assert np.array_equal(g["array"], w["array"]), (case["name"], g["node"])
            if g["op"] == "attr":
                assert g["value"] == w["value"], (case["name"], g["name"])
        out = str(tmp_path / (case["name"] + ".ref.bin"))
        ref_bin_utils.ecsave2(out, apm)
        with open(src, "rb") as a, open(out, "rb") as b:
            assert a.read() == b.read(), case["name"]


def test_non_ascii_names_are_written_as_the_reference_writes_them(tmp_path):
    """ecsave2 packs '<{len(name)}s': len(name) BYTES of the UTF-8 encoding, i.e. a non-ASCII name is cut.  Our
    writer must produce the same bytes (checked against the unmodified reference when it is here) and our
    reader must be able to read its own file back."""
    from alntools_b200 import bin_utils
    from alntools_b200.apm import ApmArrays
    haps, targets, samples = ["A", "Bé"], ["tränscript1", "t2", "ターゲット"], ["sämple"]
    lengths = np.array([[100, 101], [200, 0], [300, 301]], dtype=np.int32)
    a = (np.array([0, 2, 3], dtype=np.int32), np.array([0, 2, 1], dtype=np.int32), np.array([3, 1, 2], dtype=np.int32))
    n = (np.array([0, 2], dtype=np.int32), np.array([0, 1], dtype=np.int32), np.array([5, 7], dtype=np.int32))
    ours = str(tmp_path / "ours.bin")
    bin_utils.ecsave2_arrays(ours, haps, targets, lengths, samples, a, n)
    back = bin_utils.ecload_arrays(ours)
    assert [len(x) for x in back["targets"]] <= [len(x) for x in targets] and back["haplotypes"][0] == "A"
    assert np.array_equal(back["a"][2], a[2]) and np.array_equal(back["n"][2], n[2])
    assert np.array_equal(back["lengths"], lengths)
    from oracle import run_reference
    if not run_reference.available():
        return
    run_reference._import_reference()
    from alntools import bin_utils as ref_bin_utils
    ref = str(tmp_path / "ref.bin")
    ref_bin_utils.ecsave2(ref, ApmArrays(haps, targets, lengths.copy(), samples, a, n))
    with open(ours, "rb") as x, open(ref, "rb") as y:
        assert x.read() == y.read()


def test_c_per_cell_oracle_equals_the_python_statement():
    """oracle/ec_oracle.c: ec_oracle_build_cells (used at sizes the Python statement cannot finish) against
    oracle/ec_oracle.py: ec_from_columns_cells, which follows bam_utils_multisample.py:503-636,702-791 and is
    pinned by the multisample goldens: random files, cells, minimum counts, dropped last reads."""
    from oracle import c_oracle, ec_oracle
    from alntools_b200 import synth
    rng = np.random.default_rng(3)
    for case in range(40):
        pushes = []
        for f in range(int(rng.integers(1, 5))):
            cols = synth.make_columns(int(rng.integers(1, 200)), int(rng.integers(2, 30)), int(rng.integers(1, 4)),
                                      seed=int(rng.integers(0, 1 << 30)), mode=["light", "diploid", "heavy"][int(rng.integers(0, 3))],
                                      n_cells=int(rng.integers(1, 12)), dup_rate=0.1)
            pushes.append((cols["read_group"], cols["target_idx"], cols["hap_idx"], cols["cell_idx"], bool(rng.integers(0, 2))))
        minimum = int(rng.choice([-1, 0, 1, 2, 5, 20]))
        try:
            want = ec_oracle.ec_from_columns_cells(pushes, minimum)
        except ValueError:
            with pytest.raises(ValueError):
                c_oracle.ec_from_columns_cells(pushes, minimum)
            continue
        got = c_oracle.ec_from_columns_cells(pushes, minimum)
        for k in want:
            assert np.array_equal(want[k], got[k]), (case, k)

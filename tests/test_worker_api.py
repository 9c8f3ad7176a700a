"""The reference's chunk-worker interface (ConvertParams -> process_convert_bam ->
ConvertResults through wrapper_convert) mirrored in alntools_b200.bam_utils.  The reference plans the chunks
(its own calculate_chunks), both workers get the same ConvertParams, and the results must be equal: `ec` keys
(tids as strings, sorted as strings), their order, their counts, the alignment totals and the position
ranges.  The GPU builder is replaced by an oracle-backed stand-in (this is a CPU test of the host logic: chunk
assembly, key strings, ordering; the builder itself has its GPU parity tests)."""
import os
import sys

import numpy as np
import pytest

from conftest import GOLDEN, golden_cases

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


class _OracleBuilder(object):
    """EcBuilder stand-in: collects the pushed columns and lets the C oracle build the matrices."""

    def __init__(self, n_targets, n_haps, **kw):
        self.cols = []

    def push(self, rg, tg, hp, cell=None, order_base=0, drop_last_group=False, n=None):
        base = self.cols[-1][0][-1] + 1 - rg[0] if self.cols else 0     # keep read groups apart across pushes
        self.cols.append((np.asarray(rg) + base, np.asarray(tg), np.asarray(hp)))

    def finalize(self, *a, **k):
        from oracle import c_oracle
        rg, tg, hp = (np.concatenate([c[i] for c in self.cols]).astype(np.int32) for i in range(3))
        indptr, indices, data, counts, n_reads = c_oracle.ec_from_columns(rg, tg, hp)
        return {"n_ec": len(counts), "n_reads": n_reads, "a_indptr": indptr, "a_indices": indices, "a_data": data,
                "n_data": counts}

    def close(self):
        pass


def _reference():
    from oracle import run_reference
    if not run_reference.available():
        pytest.skip("the reference tree is not on this machine")
    bam_utils, _ = run_reference._import_reference()
    run_reference.patch_pep479(bam_utils)
    return bam_utils


@pytest.mark.parametrize("n_chunks,n_procs", [(1, 1), (3, 1), (4, 2), (8, 3)])
def test_worker_returns_what_the_reference_worker_returns(tmp_path, monkeypatch, n_chunks, n_procs):
    ref = _reference()
    from alntools_b200 import bam_utils as ours, utils
    monkeypatch.setattr(ours, "EcBuilder", _OracleBuilder)
    for case in [c for c in golden_cases("single") if c["name"] in ("synth_h2", "mixed_single", "synth_h8_heavy")]:
        bam = os.path.join(GOLDEN, case["bam"])
        chunks = ref.calculate_chunks(bam, n_chunks)
        if chunks is None:                                   # the reference's planner gives up on some splits
            continue
        for pid, ids in enumerate(utils.partition(list(range(n_chunks)), n_procs)):
            params = []
            for mod in (ref, ours):
                cp = mod.ConvertParams()
                cp.input_file, cp.temp_dir, cp.process_id, cp.track_ranges = bam, str(tmp_path), pid, True
                for cid in ids:
                    rec = chunks[cid]
                    cp.data.append((cid, rec if mod is ref else ours.ParseRecord(*rec)))
                params.append(cp)
            want = ref.wrapper_convert((params[0],))
            got = ours.wrapper_convert((params[1],))
            assert list(got.ec.items()) == list(want.ec.items()), (case["name"], n_chunks, pid)
            assert got.valid_alignments == want.valid_alignments and got.all_alignments == want.all_alignments
            assert got.tid_ranges == want.tid_ranges, (case["name"], n_chunks, pid)
            assert not os.listdir(str(tmp_path))             # temporary chunk files are gone


def test_worker_structs_have_the_reference_fields():
    ref = _reference()
    from alntools_b200 import bam_utils as ours
    assert ours.ParseRecord._fields == ref.ParseRecord._fields
    for name in ("ConvertParams", "ConvertResults"):
        assert getattr(ours, name).slots == getattr(ref, name).slots
        assert vars(getattr(ours, name)()) == vars(getattr(ref, name)())

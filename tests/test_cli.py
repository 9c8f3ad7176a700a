"""The drop-in boundary at its outermost layer: `bam2ec` / `bam2emase` keep the
reference's arguments, options and defaults (compared with the reference's own click commands when the
reference is present) and reach convert() with the reference's argument meaning.  The GPU builder is replaced
by a stub: this is about the plumbing, the matrices are the GPU tests' business."""
import os
import sys

import numpy as np
import pytest
from click.testing import CliRunner

from conftest import GOLDEN, golden_cases

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _surface(command):
    out = {}
    for p in command.params:
        out[p.name] = (type(p).__name__, tuple(sorted(p.opts)), p.default if not callable(p.default) else None,
                       getattr(p, "is_flag", False), getattr(p, "count", False), p.required)
    return out


def test_commands_have_the_reference_surface():
    from alntools_b200 import cli as ours
    sys.path.insert(0, ROOT)
    from oracle import run_reference
    if not run_reference.available():
        pytest.skip("the reference tree is not on this machine")
    run_reference._import_reference()
    from alntools import cli as ref                     # the reference package
    for name in ("bam2ec", "bam2emase", "ec2emase"):
        assert _surface(ours.cli.commands[name]) == _surface(ref.cli.commands[name]), name
    # the facade and the converters below it: same parameter names, order and defaults
    # (methods.py:32-58, bam_utils.py:512, bam_utils_multisample.py:357); ours may add trailing keywords
    import inspect
    from alntools import methods as ref_methods, bam_utils as ref_bu, bam_utils_multisample as ref_bum
    from alntools_b200 import methods as our_methods, bam_utils as our_bu, bam_utils_multisample as our_bum

    def params(fn):
        return [(p.name, p.default) for p in inspect.signature(fn).parameters.values()]
    for name in ("bam2ec", "bam2emase", "bam2ec_multisample", "bam2emase_multisample"):
        assert params(getattr(our_methods, name)) == params(getattr(ref_methods, name)), name
    for ours_fn, ref_fn in ((our_bu.convert, ref_bu.convert), (our_bum.convert, ref_bum.convert)):
        want = params(ref_fn)
        assert params(ours_fn)[:len(want)] == want, ours_fn.__module__


class _Stub(object):
    calls = []

    def __init__(self, n_targets, n_haps, with_cells=False, **kw):
        self.with_cells = with_cells
        self.rows = 0

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False

    def push(self, rg, tg, hp, cell=None, order_base=0, drop_last_group=False, n=None):
        self.rows += int(n) if n is not None else len(rg)

    def finalize(self, min_cell_count=0, copy=True):
        _Stub.calls.append(("finalize", self.with_cells, min_cell_count, self.rows))
        i32 = lambda *v: np.array(v, dtype=np.int32)
        return {"n_ec": 2, "n_reads": 3, "n_samples": 2, "a_indptr": i32(0, 1, 3), "a_indices": i32(0, 0, 1),
                "a_data": i32(1, 3, 1), "n_indptr": i32(0, 1, 2), "n_indices": i32(0, 1), "n_data": i32(2, 1),
                "cell_order": i32(0, 1)}


def test_bam2ec_command_line_reaches_convert(tmp_path, monkeypatch):
    from alntools_b200 import bam_utils, bam_utils_multisample, bin_utils, cli
    monkeypatch.setattr(bam_utils, "EcBuilder", _Stub)
    monkeypatch.setattr(bam_utils_multisample, "EcBuilder", _Stub)
    _Stub.calls = []
    runner = CliRunner()
    case = next(c for c in golden_cases("single") if c["targets"])
    bam, targets = os.path.join(GOLDEN, case["bam"]), os.path.join(GOLDEN, case["targets"])
    out, rng = str(tmp_path / "o.bin"), str(tmp_path / "r.txt")
    res = runner.invoke(cli.cli, ["bam2ec", bam, out, "-c", "3", "-p", "2", "-t", targets, "--rangefile", rng, "-v"])
    assert res.exit_code == 0, res.output + repr(res.exception)
    ec = bin_utils.ecload_arrays(out)
    assert ec["samples"] == [os.path.basename(bam)]                 # bam_utils.py:552-554
    with open(targets) as fh:
        first = [line.split()[0] for line in fh if line.strip() and not line.startswith("#")]
    assert ec["targets"][:len(first)] == first                       # target file ids come first (:571-579)
    assert os.path.getsize(rng) > 0
    # per-cell: a directory of BAM files, -m is the minimum count
    mcase = golden_cases("multisample")[0]
    mout = str(tmp_path / "m.bin")
    res = runner.invoke(cli.cli, ["bam2ec", os.path.join(GOLDEN, mcase["dir"]), mout, "--multisample", "-m", "7"])
    assert res.exit_code == 0, res.output + repr(res.exception)
    assert ("finalize", True, 7, _Stub.calls[-1][3]) == _Stub.calls[-1]
    assert len(bin_utils.ecload_arrays(mout)["samples"]) == 2
    # the reference's refusal of -s together with --multisample (cli.py:61-63)
    res = runner.invoke(cli.cli, ["bam2ec", os.path.join(GOLDEN, mcase["dir"]), mout, "--multisample", "-s", "x"])
    assert res.exit_code == 0 and "should NOT be specified" in res.output
    # a missing input is click's error, as in the reference
    res = runner.invoke(cli.cli, ["bam2ec", str(tmp_path / "missing.bam"), out])
    assert res.exit_code == 2

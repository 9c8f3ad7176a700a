import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


def golden_cases(kind=None):
    with open(os.path.join(GOLDEN, "manifest.json")) as fh:
        cases = json.load(fh)
    with open(os.path.join(GOLDEN, "manifest_range.json")) as fh:      # --rangefile goldens carry EC files too
        cases += json.load(fh)
    return [c for c in cases if kind is None or c["kind"] == kind]


@pytest.fixture(scope="session")
def built():
    """Compile libecb200.so and the C oracle once per session (nvcc cross-compiles without a GPU)."""
    import __graft_entry__
    __graft_entry__.build()
    return True


def load_records(path):
    from alntools_b200 import bam_io
    raw = bam_io.inflate_file(path)
    header = bam_io.parse_header(raw)
    return header, list(bam_io.iter_records(raw, header.records_offset))

"""Run the UNMODIFIED reference (/root/reference/alntools) through the import shims.

TEST INFRASTRUCTURE ONLY, build-container only: /root/reference does not exist on the GPU box.
Used by oracle/make_golden.py to mint tests/golden/*.bin and by tests that (when the reference is
present) cross-check the oracle restatement against the real thing.

The reference's multi-chunk path breaks on Python >= 3.7 (PEP 479: StopIteration escaping the
FastBgzfBlocks generator, bam_utils.py:1320-1343, is swallowed at :1303-1304).  `patch_pep479`
swaps in an equivalent generator FROM THE HARNESS; no reference file is edited.
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE_ROOT = os.environ.get("ALNTOOLS_REFERENCE", "/root/reference")


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "alntools"))


def _import_reference():
    shims = os.path.join(HERE, "shims")
    for p in (REFERENCE_ROOT, shims):
        if p not in sys.path:
            sys.path.insert(0, p)
    from alntools import bam_utils, bam_utils_multisample  # noqa: the reference package
    return bam_utils, bam_utils_multisample


def patch_pep479(bam_utils):
    def fast_bgzf_blocks(handle):
        data_start = 0
        while True:
            start_offset = handle.tell()
            try:
                block_length, data_len = bam_utils._quick_bgzf_load(handle)
            except StopIteration:
                return
            yield start_offset, block_length, data_start, data_len
            data_start += data_len
    bam_utils.FastBgzfBlocks = fast_bgzf_blocks


def bam2ec(bam_filename, ec_filename, num_chunks=1, number_processes=1, target_filename=None,
           temp_dir=None, range_filename=None):
    bam_utils, _ = _import_reference()
    patch_pep479(bam_utils)
    bam_utils.convert(bam_filename, ec_filename, None, num_chunks=num_chunks,
                      number_processes=number_processes,
                      temp_dir=temp_dir or os.path.dirname(ec_filename),
                      range_filename=range_filename, target_filename=target_filename)


def bam2ec_multisample(bam_dir, ec_filename, minimum_count, number_processes=1, target_filename=None,
                       range_filename=None):
    _, multi = _import_reference()
    multi.convert(bam_dir, ec_filename, None, 0, minimum_count, number_processes,
                  os.path.dirname(ec_filename), range_filename, target_filename)


if __name__ == "__main__":
    import argparse
    ap = argparse.ArgumentParser()
    ap.add_argument("bam")
    ap.add_argument("ec")
    ap.add_argument("--multisample", action="store_true")
    ap.add_argument("-c", type=int, default=1)
    ap.add_argument("-p", type=int, default=1)
    ap.add_argument("-m", type=int, default=1)
    ap.add_argument("-t", default=None)
    a = ap.parse_args()
    if a.multisample:
        bam2ec_multisample(a.bam, a.ec, a.m, a.p, a.t)
    else:
        bam2ec(a.bam, a.ec, a.c, a.p, a.t)

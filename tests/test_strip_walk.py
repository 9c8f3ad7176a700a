"""CPU check of the strip kernel's per-lane walk (alntools_b200/csrc/ecb_strip.cuh): the functions the
kernel calls for a lane are `__host__ __device__`; tests/native/strip_host_test.cu emulates every tile and
lane with them and compares the reads they close with a serial statement of the grouping rule
(alntools/bam_utils.py:301-344 on columns).  Needs nvcc (the header is CUDA source), no GPU."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_lane_walk_closes_every_read_once_with_the_right_key(tmp_path):
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    if not os.path.isfile(nvcc):
        nvcc = shutil.which("nvcc")
    if not nvcc:
        pytest.skip("nvcc not found")
    exe = str(tmp_path / "strip_host_test")
    subprocess.check_call([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-O2", "-std=c++17",
                           os.path.join(ROOT, "tests", "native", "strip_host_test.cu"), "-o", exe], cwd=ROOT)
    out = subprocess.run([exe, "600"], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout[-2000:]
    assert "600 cases, 0 failed" in out.stdout

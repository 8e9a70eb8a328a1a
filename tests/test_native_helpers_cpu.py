"""The __host__ __device__ arithmetic shared by the build and scan kernels (muscato_b200/csrc/common.cuh: fingerprints,
home buckets, both fronts' addressing, the even-bit compaction behind the exact front's member masks), checked on
the CPU: tests/native/helpers_host_test.cu is compiled by nvcc for sm_100a and its host main() run here (it makes no
CUDA runtime call, so it needs no GPU)."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shared_kernel_arithmetic_on_the_host(tmp_path):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    exe = str(tmp_path / "helpers_host_test")
    src = os.path.join(ROOT, "tests", "native", "helpers_host_test.cu")
    r = subprocess.run([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-std=c++17", "-O1", "-o", exe, src],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and r.stdout.strip().endswith("ok"), r.stdout + r.stderr

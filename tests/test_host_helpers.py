"""Host-side helpers of the CUDA sources, compiled and run on the CPU (nvcc is the only toolchain with the CUDA headers)."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(shutil.which("nvcc") is None, reason="nvcc not on PATH")
def test_fastdiv_matches_integer_division(tmp_path):
    exe = str(tmp_path / "fastdiv_check")
    src = os.path.join(ROOT, "tests", "host", "fastdiv_check.cu")
    build = subprocess.run(["nvcc", "-std=c++17", "-O2", "-o", exe, src], capture_output=True, text=True)
    assert build.returncode == 0, build.stderr[-2000:]
    run = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert run.returncode == 0 and run.stdout.strip() == "bad=0", run.stdout + run.stderr

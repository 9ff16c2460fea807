"""CPU check of the spectral C(r) kernel's per-thread steps: csrc/fft64.cuh and csrc/ising_corr_fft.cuh are host + device
code, tests/host/corr_fft_host.cu emulates the 64 threads of a set pass by pass and compares the radial sums with a
direct evaluation of the reference's loops (analysis/correlation.py:8-47) on ordered, domain, magnetised and random
lattices, with missing chains and in the two-groups form.  Needs nvcc (host compilation only), no GPU."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(shutil.which("nvcc") is None, reason="nvcc not on PATH")
def test_spectral_correlation_steps_on_the_host(tmp_path):
    exe = str(tmp_path / "corr_fft_host")
    src = os.path.join(ROOT, "tests", "host", "corr_fft_host.cu")
    subprocess.check_call(["nvcc", "-O2", "-std=c++17", "-Wno-deprecated-gpu-targets", "-o", exe, src])
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "mismatches 0" in r.stdout

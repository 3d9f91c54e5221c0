"""The C++ mirror of the reference's module API (host/fcp_host.hpp) driving the C-ABI: the Poisson application of
applications/Poisson/poisson.f90.  CPU: it builds, links against the library and refuses to run without a GPU.
GPU: known answer + the reference's report lines."""
import os
import re
import subprocess

import numpy as np
import pytest

from conftest import EMU, HAS_GPU

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
APP = os.path.join(ROOT, "host", "poisson_app")


def build():
    # FCP_TEST_EMU=1 (tests/emu_hook.py): link the host program against the emulation build of the same sources
    extra = ["LIBDIR=../tests/emu", "LIB=fcp_emu"] if EMU else []
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "host"), "-B", "poisson_app"] + extra, stdout=subprocess.DEVNULL)


@pytest.mark.skipif(HAS_GPU, reason="checks the no-GPU failure mode")
def test_builds_and_fails_loudly_without_gpu():
    build()
    r = subprocess.run([APP, "8"], capture_output=True, text=True)
    assert r.returncode != 0
    assert "no CUDA device" in r.stderr or "code -3" in r.stderr or "code -2" in r.stderr


@pytest.mark.gpu
def test_poisson_app_known_answer(orc):
    build()
    errs = {}
    for n in (16, 32, 64):
        r = subprocess.run([APP, str(n)], capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stderr
        lines = r.stdout.splitlines()
        assert re.match(r"  PCG\(IC0\):  Solving for p, Initial residual = +\d\.\d{3}E[+-]\d\d, Final residual = +\d\.\d{3}E[+-]\d\d, No Iterations \d+$", lines[0]), lines[0]
        assert lines[2].startswith("  PCG(Jacobi):  Solving for p,")
        m = re.search(r"POISSON_APP_DONE (\d+) (\S+) (\S+)", r.stdout)
        errs[n] = (float(m.group(2)), float(m.group(3)))
        assert abs(errs[n][0] - errs[n][1]) < 1e-9       # both solvers reach the same discrete solution
    # second order in h (applications/Poisson/poisson.f90:104 prints exactly these two numbers for a grid-convergence check)
    assert errs[32][0] < errs[16][0] / 3.5 and errs[64][0] < errs[32][0] / 3.5, errs

"""CPU: the C-ABI library loads, exports every symbol include/fcp.h declares, and refuses to run without a GPU
(no CPU fallback).  No compute calls here."""
import ctypes
import os

import numpy as np
import pytest

from fcb200 import lib as L
from fcb200 import mesh as M
from conftest import HAS_GPU


def test_library_exports_every_declared_symbol():
    assert os.path.exists(L.SO_PATH), "csrc/libfcp_b200.so not built: run __graft_entry__.build()"
    so = ctypes.CDLL(L.SO_PATH)
    names = L.declared_symbols()
    assert len(names) >= 30
    missing = [n for n in names if not hasattr(so, n)]
    assert not missing, missing


def test_version_and_error_string():
    so = L.lib()
    assert so.fcp_version() >= 100
    assert isinstance(so.fcp_last_error(), bytes)
    assert L.launch_count() >= 0


def test_field_table_matches_header():
    import re
    with open(L.HEADER) as fh:
        text = fh.read()
    block = text[text.index("FCP_F_U = 0"): text.index("FCP_F_COUNT")]
    names = re.findall(r"FCP_F_([A-Z0-9]+)", block)
    assert names == L.FIELDS


@pytest.mark.skipif(HAS_GPU, reason="checks the no-GPU failure mode")
def test_no_cpu_fallback_without_gpu():
    m = M.cavity_mesh(3)
    with pytest.raises(L.FcpError) as ei:
        L.Context(m)
    assert "no CUDA device" in str(ei.value) or "code -3" in str(ei.value) or "code -2" in str(ei.value)
    with pytest.raises(L.FcpError):
        L.CsrSolver(np.array([1, 2], np.int32), np.array([1], np.int32), np.array([1], np.int32))


def test_report_line_format():
    """linear_solvers.f90:354-355: '(3a,1PE10.3,a,1PE10.3,a,I0)'"""
    rep = L.Report(res0=1.0, resl=2.5e-9, factor=4.0, resor=0.25, iters=17, solver=L.SOLVER_ICCG)
    assert L.report_line(rep, "p") == "  PCG(IC0):  Solving for p, Initial residual =  2.500E-01, Final residual =  6.250E-10, No Iterations 17"
    rep = L.Report(res0=3.0e-15, resl=3.0e-15, factor=0.0, resor=3.0e-15, iters=0, solver=L.SOLVER_DPCG)
    assert L.report_line(rep, "U") == "  PCG(Jacobi):  Solving for U, Initial residual =  3.000E-15, Final residual =  3.000E-15, No Iterations 0"

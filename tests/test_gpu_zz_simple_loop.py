"""GPU (B200): whole SIMPLE outer iterations with every field resident on the device -- `fcp_calcuvw` followed by
`fcp_calcp_simple`, nothing uploaded in between -- on the reference's examples/cavity case.
(Written at the end of round 1 after the GPU budget was spent: both entry points are individually bit-exact on hardware
(tests/test_gpu_parity.py, tests/test_gpu_rows2.py); this file, which chains them, sorts last on purpose.)"""
import numpy as np
import pytest

import simple_loop as S
from fcb200 import lib as L

pytestmark = pytest.mark.gpu


def gpu_state(m, f):
    ctx = L.Context(m)
    for k in ("u", "v", "w", "p", "pp", "den", "vis", "apu", "apv", "apw"):
        ctx.upload(k.upper(), f[k])
    visw = np.zeros(m.numTotal); visw[m.numCells:] = f["visw"]
    ctx.upload("VISW", visw); ctx.upload("FLMASS", f["flmass"]); ctx.upload("A", np.zeros(ctx.nnz))
    return ctx


def gpu_iteration(ctx):
    I = S.INPUT
    ur = ctx.calcuvw(solver=I["lSolverU"], maxiter=I["maxiterU"], tol_abs=I["tolAbsU"], tol_rel=I["tolRelU"], urf=I["urfU"], gds=I["gdsU"],
                     cscheme=I["cSchemeU"], pscheme=I["pscheme"], viscos=I["viscos"])
    pr = ctx.calcp_simple(solver=I["lSolverP"], maxiter=I["maxiterP"], tol_abs=I["tolAbsP"], tol_rel=I["tolRelP"], urfp=I["urfP"], npcor=1,
                          pRefCell=I["pRefCell"], pscheme=I["pscheme"], const_mflux=True)
    return ur, pr[0]


def test_simple_iterations_bit_identical_to_oracle(fcp, orc):
    """8 outer iterations: u, v, w, p, pp, flmass, apu, the matrix and every solver count equal the oracle's (TREE sums)."""
    m = S.cavity(39)
    c = orc.Csr(m)
    f = S.initial_state(m)
    ctx = gpu_state(m, f)
    prm = S.oracle_params(orc, orc.SUM_TREE)
    a = np.zeros(c.nnz)
    dP = np.zeros((m.numTotal, 3))
    for it in range(8):
        ur, pr = gpu_iteration(ctx)
        our, opr = S.oracle_iteration(orc, m, c, prm, f, a, dP, orc.SUM_TREE)
        assert [r.iters for r in ur] == [r.iters for r in our] and pr.iters == opr.iters, (it, [r.iters for r in ur], [r.iters for r in our], pr.iters, opr.iters)
        for k in ("u", "v", "w", "p", "pp"):
            g = ctx.download(k.upper())
            assert np.array_equal(g, f[k]), (it, k, float(np.abs(g - f[k]).max()))
        assert np.array_equal(ctx.download("FLMASS"), f["flmass"]), (it, "flmass")
        assert np.array_equal(ctx.download("A"), a), (it, "a")
    ctx.close()


def test_cavity_re100_on_gpu_matches_ghia(fcp):
    """400 device-resident SIMPLE iterations reproduce Ghia, Ghia & Shin (1982), Re = 100, like the CPU oracle does."""
    m = S.cavity(39)
    ctx = gpu_state(m, S.initial_state(m))
    first = last = None
    for it in range(400):
        ur, pr = gpu_iteration(ctx)
        first = first or ur[0].resor
        last = ur[0].resor
    assert last < 1e-4 * first
    y, u = S.centreline_u(m, ctx.download("U"))
    err = np.abs(np.interp(S.GHIA_Y, y, u) - S.GHIA_U)
    assert err.max() < 0.005, err.max()
    ctx.close()

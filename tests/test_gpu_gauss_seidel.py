"""GPU (B200): the 'gauss-seidel' branch of csrsolve (linear_solvers.f90:63-66, 96-201) -- the level-scheduled sweep with its separate
new-value array must give the sequential reference sweep's bits: solution, residual norms, sweep count and the report line, on structured,
distorted and polyhedral patterns, on a periodic one (twin entries), on a structurally NONSYMMETRIC pattern (where a row of an earlier level
can have a higher index than a row that still needs its old value), and through the reference's quirky early return (fi is updated before
the `res0 < tol_abs` test)."""
import json
import os

import numpy as np
import pytest

import cases
from fcb200 import lib as L
from fcb200 import mesh as M

pytestmark = pytest.mark.gpu


def _same(rep, ro):
    assert rep.iters == ro.iters and rep.res0 == ro.res0 and rep.resl == ro.resl and rep.factor == ro.factor and rep.resor == ro.resor, \
        (rep.as_dict(), (ro.iters, ro.res0, ro.resl, ro.factor, ro.resor))


@pytest.mark.parametrize("name", ["cavity8", "distorted", "poly", "channel_periodic"])
def test_gauss_seidel_on_the_pressure_system_matches_the_oracle(fcp, orc, name):
    ms = cases.meshes()
    m = ms[name] if name in ms else {"cavity8": M.cavity_mesh(8), "distorted": M.cavity_mesh(7, distort=0.25), "poly": M.polyhedral_mesh(6),
                                     "channel_periodic": cases.periodic_channel()}[name]
    c = orc.Csr(m)
    f = cases.fields(m)
    ctx = L.Context(m)
    for k, v in f.items():
        ctx.upload(k.upper(), v)
    ctx.gradp_and_sources("linear", "P")
    ctx.assemble_pcorr_simple()
    a, su = ctx.download("A"), ctx.download("SU")[: m.numCells]
    for itr_max, tol_rel in ((7, 1e-30), (400, 1e-2), (0, 1e-2)):
        ctx.upload("PP", f["pp"])
        rep = ctx.csrsolve("gauss-seidel", "PP", "SU", itr_max, 1e-30, tol_rel)
        x = f["pp"].copy()
        ro = orc.solve(orc.GAUSS_SEIDEL, c.ia, c.ja, a, c.diag, x, su, itr_max, 1e-30, tol_rel, orc.SUM_TREE)
        _same(rep, ro)
        assert np.array_equal(ctx.download("PP")[: m.numCells], x[: m.numCells]), (name, itr_max)
        assert L.report_line(rep, "p") == orc.report_line(orc.GAUSS_SEIDEL, "p", ro)
        if itr_max == 400:
            assert 1 < rep.iters < 400 and rep.resl < 1e-2 * rep.res0
    ctx.close()


def test_gauss_seidel_early_return_updates_fi_first(fcp, orc):
    m = M.cavity_mesh(6, distort=0.1)
    c, a, su = cases.poisson_system(m, orc)
    s = L.CsrSolver(c.ia, c.ja, c.diag)
    x = np.full(m.numCells, 0.0); xo = x.copy()
    rep = s.solve("gauss-seidel", a, x, su, 50, 1e30, 1e-9)          # tol_abs huge: "Initial residual = res0 ... No Iterations 1"
    ro = orc.solve(orc.GAUSS_SEIDEL, c.ia, c.ja, a, c.diag, xo, su, 50, 1e30, 1e-9, orc.SUM_TREE)
    _same(rep, ro)
    assert rep.iters == 1 and rep.factor == 0.0 and np.abs(x).max() > 0.0 and np.array_equal(x, xo)
    assert L.report_line(rep, "T").endswith("No Iterations 1") and L.report_line(rep, "T") == orc.report_line(orc.GAUSS_SEIDEL, "T", ro)
    s.close()


def test_gauss_seidel_structurally_nonsymmetric_pattern(fcp, orc):
    """Rows with no lower entries sit in level 0 whatever their index; a lower-index row of a later level that has such a row as an UPPER
    neighbour must still read its old value."""
    rng = np.random.default_rng(11)
    n = 300
    rows = []
    for i in range(n):
        cols = {i}
        if i % 3 != 0:                                   # every third row has no lower entries at all
            cols |= set(rng.choice(i, size=min(i, 3), replace=False).tolist()) if i else set()
        cols |= set(rng.choice(np.arange(i + 1, n), size=min(n - 1 - i, 3), replace=False).tolist()) if i < n - 1 else set()
        rows.append(sorted(cols))
    ia = np.ones(n + 1, np.int32)
    ja, av, diag = [], [], np.zeros(n, np.int32)
    for i, cols in enumerate(rows):
        ia[i + 1] = ia[i] + len(cols)
        off = -rng.random(len(cols))
        for k, cj in enumerate(cols):
            if cj == i:
                diag[i] = len(ja) + 1
                av.append(1.0 + len(cols))
            else:
                av.append(off[k])
            ja.append(cj + 1)
    ja, av = np.array(ja, np.int32), np.array(av)
    b = rng.standard_normal(n)
    s = L.CsrSolver(ia, ja, diag)
    for itr_max, tol_rel in ((5, 1e-30), (200, 1e-10)):
        x = np.zeros(n); xo = np.zeros(n)
        rep = s.solve("gauss-seidel", av, x, b, itr_max, 1e-30, tol_rel)
        ro = orc.solve(orc.GAUSS_SEIDEL, ia, ja, av, diag, xo, b, itr_max, 1e-30, tol_rel, orc.SUM_TREE)
        _same(rep, ro)
        assert np.array_equal(x, xo)
    import scipy.sparse as sp
    A = sp.csr_matrix((av, ja - 1, ia - 1), shape=(n, n))
    assert np.abs(A @ x - b).sum() <= 1e-9 * np.abs(b).sum()
    s.close()


def test_gauss_seidel_golden_5x5(fcp):
    """the SPD system of test/test_linear_solvers_spsolve.f90:129-138 (known answer to two decimals)"""
    with open(os.path.join(cases.GOLDEN, "spsolve_5x5.json")) as fh:
        g = json.load(fh)
    s = L.CsrSolver(g["ioffset"], g["ja"], g["diag"])
    x = np.zeros(5)
    rep = s.solve("gauss-seidel", g["spd"]["a_f32"], x, g["spd"]["b_f32"], 2000, 1e-12, 1e-10)
    assert rep.iters < 2000
    np.testing.assert_allclose(x, g["spd"]["x"], atol=0.0051 + 2e-3 * np.abs(g["spd"]["x"]).max())
    s.close()

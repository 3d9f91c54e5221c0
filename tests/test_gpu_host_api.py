"""GPU: the reference's module-level API (freecappuccino-dev_b200/host.py) used the way the reference's own drivers use it."""
import io
import re

import numpy as np
import pytest

import cases
from fcb200 import host as H
from fcb200 import mesh as M

pytestmark = pytest.mark.gpu


def test_wall_distance_like_the_reference(orc):
    """src/mesh/wall_distance.f90:96-133: laplacian(1,phi) ; csrsolve('iccg', 500, 1e-12, 1e-10) ; grad_gauss ; d = -|g| + sqrt(|g|^2 + 2 phi)."""
    m = M.cavity_mesh(24)
    out = io.StringIO()
    case = H.Case(m, out=out)
    n = m.numCells
    case.su[:] = -m.vol[:n]
    phi = np.zeros(m.numTotal)
    case.laplacian(np.ones(m.numTotal), phi)
    res0 = case.csrsolve("iccg", phi, case.su, 500, 1e-12, 1e-10, "wdis")
    g = np.zeros((m.numTotal, 3))
    case.grad_gauss(phi, g)
    gm = np.sqrt((g[:n] ** 2).sum(1))
    d = -gm + np.sqrt(gm * gm + 2 * phi[:n])
    exact = np.minimum.reduce([m.xc[:n], 1 - m.xc[:n], m.yc[:n], 1 - m.yc[:n], m.zc[:n], 1 - m.zc[:n]])
    near = exact < 0.1
    assert np.abs(d[near] - exact[near]).max() < 0.03
    line = out.getvalue().rstrip("\n")
    assert re.match(r"  PCG\(IC0\):  Solving for wdis, Initial residual = +\S+, Final residual = +\S+, No Iterations \d+$", line), line
    # same numbers as the oracle driven the same way
    c = orc.Csr(m)
    su = -m.vol[:n].copy()
    a = orc.laplacian(m, c, np.ones(m.numTotal), np.zeros(m.numTotal), su)
    x = np.zeros(n)
    rep = orc.solve(orc.ICCG, c.ia, c.ja, a, c.diag, x, su, 500, 1e-12, 1e-10, orc.SUM_TREE)
    assert np.array_equal(case.a, a) and np.array_equal(phi[:n], x) and res0 == rep.resor
    assert line == orc.report_line(orc.ICCG, "wdis", rep)
    case.close()


def test_calcp_simple_module_api(orc):
    m = cases.meshes()["hex10_distorted"]
    f = cases.fields(m)
    case = H.Case(m, out=io.StringIO())
    for k in ("u", "v", "w", "p", "den", "apu", "apv", "apw"):
        getattr(case, k)[:] = f[k]
    case.gradp_and_sources(case.p)
    case.lSolverP, case.maxiterP, case.tolRelP, case.urfP = "iccg", 20, 0.025, 0.3       # examples/cavity/input.nml settings
    reps = case.calcp_simple()
    assert 0 < reps[0].iters <= 20
    g = {k: v.copy() for k, v in f.items()}
    c = orc.Csr(m)
    dP = np.zeros((m.numTotal, 3))
    orc.gradp_and_sources(m, 0, g["p"], g["apu"], dP)
    g["pp"][:] = 0
    a, su, flm = orc.assemble_pcorr(m, c, g["den"], g["u"], g["v"], g["w"], g["p"], g["pp"], dP, g["apu"])
    rep = orc.solve(orc.ICCG, c.ia, c.ja, a, c.diag, g["pp"], su, 20, 1e-13, 0.025, orc.SUM_TREE)
    orc.correct_simple(m, c, 0, a, g["den"], g["u"], g["v"], g["w"], g["p"], g["pp"], g["apu"], g["apv"], g["apw"], 0.3, 1, dP, flm)
    assert reps[0].iters == rep.iters
    for k in ("u", "v", "w", "p", "pp"):
        assert np.array_equal(getattr(case, k), g[k]), k
    assert np.array_equal(case.flmass, flm) and np.array_equal(case.a, a)
    with pytest.raises(Exception):
        case.csrsolve("gauss-seidel", case.pp, case.su, 1, 0.0, 0.1, "p")
    case.close()

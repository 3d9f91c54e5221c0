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
        case.csrsolve("pmgmres", case.pp, case.su, 1, 0.0, 0.1, "p")       # not on the accelerated path
    # 'gauss-seidel' is: two sweeps through the module API against the oracle
    x = g["pp"].copy()
    orc.solve(orc.GAUSS_SEIDEL, c.ia, c.ja, case.a, c.diag, x, su, 2, 0.0, 1e-30, orc.SUM_TREE)
    case.csrsolve("gauss-seidel", case.pp, su, 2, 0.0, 1e-30, "p")
    assert np.array_equal(case.pp[: m.numCells], x[: m.numCells])
    case.close()


def test_turbulence_procedures_of_the_module_api(orc):
    """modify_viscosity_k_epsilon_rlzb / _k_omega_sst / _sgs, wall_distance, updateBoundary and constant_mass_flow_forcing called the way the
    reference's main loop calls them (no arguments, module state): same state as driving the C-ABI directly (tests/test_gpu_scalar.py)."""
    import test_gpu_scalar as T
    m = cases.meshes()["channel_inout"]
    g = T.scalar_inputs(m, orc)
    out = io.StringIO()
    case = H.Case(m, out=out)
    n = m.numCells
    for k in ("u", "v", "w", "den", "vis", "te", "ed", "flmass"):
        getattr(case, k)[...] = g[k]
    wall = np.concatenate([m.patch_faces(ib) - m.numInnerFaces for ib in range(m.numBoundaries) if m.bctype[ib] == M.BC_WALL])
    case.visw[...] = g["visw"][wall]; case.dnw[...] = g["dnw"][wall]
    case.viscos, case.urfVis = 0.01, 0.6
    case.TurbModelScalar = [dict(maxiter=8, tolRel=1e-4, urf=0.7, gds=0.8, cScheme="muscl")] * 2
    case.modify_viscosity_k_epsilon_rlzb()
    text = out.getvalue()
    assert re.search(r"<= k <=", text) and re.search(r"<= epsilon <=", text) and text.count("Solving for") == 2
    # the same sequence through the oracle
    c = orc.Csr(m)
    f = {k: (v.copy() if isinstance(v, np.ndarray) else v) for k, v in g.items()}
    gU, gV, gW = orc.grad_gauss(m, f["u"]), orc.grad_gauss(m, f["v"]), orc.grad_gauss(m, f["w"])
    f["magStrain"], _ = orc.calc_strain_and_vorticity(m, gU, gV, gW)
    prm = T.oracle_params(orc, orc.SC_TKE_RLZB, "bicgstab", "muscl", "gauss", "none", "steady")
    prm.prtr = 1.0
    orc.calcsc(m, c, prm, f)
    prm.kind, prm.prtr = orc.SC_EPS_RLZB, 1.0 / 1.2
    orc.calcsc(m, c, prm, f)
    orc.modify_mu_eff_rlzb(m, 0.6, 0.01, gU, gV, gW, f["te"], f["ed"], f["den"], f["u"], f["v"], f["w"], f["dnw"], f["vis"], f["visw"])
    T.close(case.te, f["te"], "te", 1e-12); T.close(case.ed, f["ed"], "ed", 1e-10); T.close(case.vis, f["vis"], "vis", 1e-10)
    T.close(case.visw, f["visw"][wall], "visw", 1e-10)
    # wall distance, SST, SGS, forcing run through and keep the fields finite and positive
    wd = case.wall_distance()
    assert wd.min() > 0 and "Solving for Wdis" in out.getvalue()
    case.ed[...] = 40.0 * case.ed
    case.modify_viscosity_k_omega_sst()
    case.modify_viscosity_sgs("wale")
    case.apu[...] = 1.0
    case.magUbar = 0.5
    case.constant_mass_flow_forcing()
    assert abs((m.vol[:n] * case.u[:n]).sum() / m.vol[:n].sum() - 0.5) < 1e-12 and "Uncorrected Ubar" in out.getvalue()
    phi = np.random.default_rng(2).standard_normal(m.numTotal)
    ref = phi.copy()
    orc.update_boundary(m, ref)
    case.updateBoundary(phi)
    assert np.array_equal(phi, ref) and np.all(np.isfinite(case.vis)) and case.vis[:n].min() > 0
    case.close()


def test_fv_equation_operators_and_axpby():
    """Row a2: type(fvEquation) with operator(+), operator(-), operator(==) (fvImplicit/fvEquation.f90:158-404, src-par/fv_equation.f90:48-71).
    equation +/- equation: coef and source element-wise (add_fvEquations / subtract_fvEquations); equation +/- source field: a NEW equation whose
    coef is zero and whose source is su +/- field (add_source_to_fvEquation :168-190, subtract_source_from_fvEquation :224-246); operator(==) is
    bound to the subtract procedures (:68-73).  One add or subtract per element has a single correctly rounded result, so the comparison with numpy
    is bit for bit; fcp_field_axpby itself is also checked with general alpha / beta (two roundings, no FMA: -fmad=false) and for its extent check."""
    from fcb200 import lib as L
    m = cases.meshes()["poly_10faces"]
    case = H.Case(m, out=io.StringIO())
    n, nnz = m.numCells, case.nnz
    rng = np.random.default_rng(2024)
    e1 = H.FvEquation(case, rng.standard_normal(nnz), rng.standard_normal(n))
    e2 = H.FvEquation(case, rng.standard_normal(nnz) * 1e3, rng.standard_normal(n) * 1e-3)
    s = e1 + e2
    assert np.array_equal(s.coef, e1.coef + e2.coef) and np.array_equal(s.source, e1.source + e2.source)
    d = e1 - e2
    assert np.array_equal(d.coef, e1.coef - e2.coef) and np.array_equal(d.source, e1.source - e2.source)
    q = e1.equals(e2)
    assert np.array_equal(q.coef, d.coef) and np.array_equal(q.source, d.source)
    src = rng.standard_normal(m.numTotal)                   # a volScalarField: numTotal values, the cells' part is used
    p = e1 + src
    assert not p.coef.any() and np.array_equal(p.source, e1.source + src[:n])
    r = e1 - src
    assert not r.coef.any() and np.array_equal(r.source, e1.source - src[:n])
    assert np.array_equal(e1.equals(src).source, r.source)
    # the inputs are untouched (the operators return new equations)
    assert e1.coef.shape == (nnz,) and e1.source.shape == (n,)
    # fcp_field_axpby, general coefficients: dst = alpha x + beta y in the matrix slots (CSR order in, CSR order out) and in the field slots
    c = case.ctx
    x, y = rng.standard_normal(nnz), rng.standard_normal(nnz)
    c.upload("A", x); c.upload("H", y)
    c.axpby("H", 0.75, "A", -2.5, "H")                      # in place in y
    assert np.array_equal(c.download("H"), 0.75 * x + (-2.5) * y)
    assert np.array_equal(c.download("A"), x)
    xs, ys = rng.standard_normal(m.numTotal), rng.standard_normal(m.numTotal)
    c.upload("S0", xs); c.upload("S1", ys)
    c.axpby("S2", 3.0, "S0", 1.0, "S1")
    assert np.array_equal(c.download("S2"), 3.0 * xs + ys)
    with pytest.raises(L.FcpError):
        c.axpby("A", 1.0, "A", 1.0, "S0")                   # extents differ: refused, nothing is written
    assert np.array_equal(c.download("A"), x)
    case.close()

"""GPU (B200): the CUDA path, called through the C-ABI (include/fcp.h via ctypes), against the CPU oracle on the same
seeded inputs.  fp64 everywhere.  Because the kernels keep the reference's per-cell / per-row accumulation order, use
no FMA contraction, and the reductions follow the fixed tree the oracle mirrors (SUM_TREE), the bar for everything
below is BIT-EXACT equality (stricter than north_star's 1e-12 relative) unless a test says otherwise."""
import json
import os

import numpy as np
import pytest

import cases
from fcb200 import lib as L
from fcb200 import mesh as M

pytestmark = pytest.mark.gpu

MESHES = ["ref400", "hex6", "hex12_graded", "hex10_distorted", "slab39_empty", "channel_inout", "channel_pressure", "poly_10faces", "hex_many_faces", "tiny3",
          "channel_periodic", "duct_periodic_x", "duct_periodic_first"]


@pytest.fixture(scope="module")
def allmeshes():
    return cases.meshes()


def make_ctx(m, f=None):
    ctx = L.Context(m)
    if f:
        for k, v in f.items():
            ctx.upload(k.upper(), v)
    return ctx


def eq(a, b, what=""):
    a = np.asarray(a); b = np.asarray(b)
    if not np.array_equal(a, b):
        d = np.abs(a - b)
        i = int(np.argmax(d))
        raise AssertionError(f"{what}: not bit-identical: max |diff| {d.max():.3e} at {i} (gpu {a.ravel()[i]!r} oracle {b.ravel()[i]!r}), "
                             f"{(a != b).sum()} of {a.size} differ")


@pytest.mark.parametrize("name", MESHES)
def test_csr_pattern(fcp, orc, allmeshes, name):
    """create_CSR_matrix, sparse_matrix.f90:86-296: ia, ja, diag and both face->slot maps identical."""
    m = allmeshes[name]
    ctx = make_ctx(m)
    ia, ja, diag, kpn, knp = ctx.csr_pattern()
    c = orc.Csr(m)
    for a, b, w in ((ia, c.ia, "ia"), (ja, c.ja, "ja"), (diag, c.diag, "diag"), (kpn, c.icell_jcell, "icell_jcell"), (knp, c.jcell_icell, "jcell_icell")):
        eq(a, b, w)
    ctx.close()


@pytest.mark.parametrize("name", MESHES)
def test_spmv_and_matrix_roundtrip(fcp, orc, allmeshes, name):
    m = allmeshes[name]
    ctx = make_ctx(m)
    c = orc.Csr(m)
    rng = np.random.default_rng(7)
    a = rng.standard_normal(c.nnz)
    x = rng.standard_normal(m.numTotal)
    ctx.upload("A", a)
    eq(ctx.download("A"), a, "CSR->SELL->CSR round trip")
    ctx.upload("S0", x)
    ctx.spmv("S0", "S1")
    eq(ctx.download("S1")[: m.numCells], orc.spmv(c.ia, c.ja, a, x), "spmv")
    ctx.close()


@pytest.mark.parametrize("name", MESHES)
def test_gradients(fcp, orc, allmeshes, name):
    """grad_gauss gradients.f90:1607, grad_lsq :782, grad_lsq_dm :1334 (+ create_matrix_*), quirks Q1/Q2 reproduced."""
    m = allmeshes[name]
    f = cases.fields(m)
    ctx = make_ctx(m)
    ctx.upload("S0", f["p"])
    ctx.grad(L.GRAD_GAUSS, "S0", "G0")
    eq(ctx.download("G0"), orc.grad_gauss(m, f["p"]), "grad_gauss")
    for meth, w in ((L.GRAD_LSQ, False), (L.GRAD_LSQ_DM, True)):
        ctx.create_lsq_grad_matrix(meth)
        D = orc.create_matrix_lsq(m, w)
        for ref_row2 in (True, False):
            ctx.grad(meth, "S0", "G0", lsq_row2_reference=ref_row2)
            eq(ctx.download("G0"), orc.grad_lsq(m, w, D, f["p"], row2_correct=not ref_row2), f"grad_lsq w={w} ref_row2={ref_row2}")
    ctx.close()


def test_gauss_linear_field_golden(fcp):
    """The reference's own gradient test (testFieldOperations.f90:137-160) on its own mesh: (1,1,1)."""
    m = cases.golden_mesh()
    ctx = make_ctx(m)
    ctx.upload("S0", m.boundary_values_of(lambda x, y, z: x + y + z))
    ctx.grad(L.GRAD_GAUSS, "S0", "G0")
    g = ctx.download("G0")
    np.testing.assert_allclose(g[: m.numCells], 1.0, rtol=0, atol=2e-12)
    pins = np.load(os.path.join(cases.GOLDEN, "oracle_pins.npz"))
    ctx.upload("S0", pins["phi"])
    ctx.grad(L.GRAD_GAUSS, "S0", "G0")
    eq(ctx.download("G0"), pins["grad_gauss"], "golden grad_gauss pin")
    ctx.close()


@pytest.mark.parametrize("name", MESHES)
def test_laplacian(fcp, orc, allmeshes, name):
    m = allmeshes[name]
    f = cases.fields(m)
    mu = np.abs(f["den"])
    c = orc.Csr(m)
    su0 = np.random.default_rng(3).standard_normal(m.numCells)
    su = su0.copy()
    a = orc.laplacian(m, c, mu, f["p"], su)
    ctx = make_ctx(m)
    ctx.upload("S0", mu); ctx.upload("S1", f["p"]); ctx.upload("SU", su0)
    ctx.laplacian("S0", "S1")
    eq(ctx.download("A"), a, "laplacian a")
    eq(ctx.download("SU")[: m.numCells], su, "laplacian su")
    ctx.close()


@pytest.mark.parametrize("name", MESHES)
@pytest.mark.parametrize("pscheme", ["linear", "central", "weighted"])
def test_gradp_and_sources(fcp, orc, allmeshes, name, pscheme):
    """Pressure/nablap.f90:19-208 + bpres.f90, all three pschemes."""
    m = allmeshes[name]
    f = cases.fields(m)
    p = f["p"].copy()
    dP = np.zeros((m.numTotal, 3))
    su, sv, sw = orc.gradp_and_sources(m, L.PSCHEME[pscheme], p, f["apu"], dP)
    ctx = make_ctx(m, dict(p=f["p"], apu=f["apu"]))
    ctx.gradp_and_sources(pscheme, "P")
    eq(ctx.download("SU")[: m.numCells], su, "su"); eq(ctx.download("SV")[: m.numCells], sv, "sv"); eq(ctx.download("SW")[: m.numCells], sw, "sw")
    eq(ctx.download("DPDXI")[: m.numCells], dP[: m.numCells], "dPdxi")
    eq(ctx.download("P"), p, "p incl. extrapolated boundary values")
    ctx.close()


def simple_oracle(orc, m, f, solver, maxiter, tol_rel, urfp=0.3, pref=1, pscheme=0, flomas=1.0, npcor=1):
    c = orc.Csr(m)
    g = {k: v.copy() for k, v in f.items()}
    dP = np.zeros((m.numTotal, 3))
    orc.gradp_and_sources(m, pscheme, g["p"], g["apu"], dP)       # what calcuvw leaves behind (velocity.f90:173)
    a = np.zeros(c.nnz); su = np.zeros(m.numCells); flm = np.zeros(m.numFaces)
    # inlet fluxes are prescribed by the host
    Fi = m.numInnerFaces
    for ib in range(m.numBoundaries):
        if m.bctype[ib] == M.BC_INLET:
            pf = m.patch_faces(ib)
            ijb = m.numCells + pf - Fi
            flm[pf] = g["den"][ijb] * (g["u"][ijb] * m.arx[pf] + g["v"][ijb] * m.ary[pf] + g["w"][ijb] * m.arz[pf])
    flm0 = flm.copy()
    orc.assemble_pcorr_into(m, c, g["den"], g["u"], g["v"], g["w"], g["p"], g["pp"], dP, g["apu"], a, su, flm, flomas=flomas, apv=g["apv"], apw=g["apw"])
    out = dict(a=a.copy(), su_asm=su.copy(), flm_asm=flm.copy(), flm0=flm0, dP0=dP.copy(), p0=g["p"].copy())
    reps = []
    for ip in range(npcor):
        rep = orc.solve(solver, c.ia, c.ja, a, c.diag, g["pp"], su, maxiter, 1e-30, tol_rel, orc.SUM_TREE)
        reps.append(rep)
        s3 = orc.correct_simple(m, c, pscheme, a, g["den"], g["u"], g["v"], g["w"], g["p"], g["pp"], g["apu"], g["apv"], g["apw"], urfp, pref, dP, flm)
        if ip != npcor - 1:
            orc.nonorth_corrector(m, g["den"], g["apu"], dP, su, flm)
    out.update(g)
    out.update(su=s3[0], sv=s3[1], sw=s3[2], dP=dP, flm=flm, reps=reps)
    return out


@pytest.mark.parametrize("name", MESHES)
def test_assemble_pcorr(fcp, orc, allmeshes, name):
    """calcp_simple.f90:69-234 + facefluxmass2 (faceflux_mass.f90:175-249) + patch terms + adjustMassFlow."""
    m = allmeshes[name]
    f = cases.fields(m)
    o = simple_oracle(orc, m, f, orc.DPCG, 0, 1.0)
    ctx = make_ctx(m, f)
    ctx.upload("FLMASS", o["flm0"])
    ctx.gradp_and_sources("linear", "P")
    ctx.assemble_pcorr_simple(False, 1.0)
    eq(ctx.download("A"), o["a"], "a")
    eq(ctx.download("SU")[: m.numCells], o["su_asm"], "su")
    eq(ctx.download("FLMASS"), o["flm_asm"], "flmass")
    ctx.close()


@pytest.mark.parametrize("name", MESHES)
@pytest.mark.parametrize("solver", ["dpcg", "iccg", "bicgstab"])
def test_calcp_simple_end_to_end(fcp, orc, allmeshes, name, solver):
    """One whole calcp_simple (assembly, solve, corrections) with 2 pressure correctors: every output field bit-identical,
    iteration counts identical."""
    m = allmeshes[name]
    f = cases.fields(m)
    sid = L.SOLVER_ID[solver]
    o = simple_oracle(orc, m, f, sid, 60, 1e-6, urfp=0.3, pref=2, npcor=2)
    ctx = make_ctx(m, f)
    ctx.upload("FLMASS", o["flm0"])
    ctx.gradp_and_sources("linear", "P")
    reps = ctx.calcp_simple(solver=solver, maxiter=60, tol_abs=1e-30, tol_rel=1e-6, urfp=0.3, npcor=2, pRefCell=2, flomas=1.0)
    for r, ro in zip(reps, o["reps"]):
        assert r.iters == ro.iters, (r.iters, ro.iters)
        assert r.res0 == ro.res0 and r.resl == ro.resl and r.resor == ro.resor
    for k in ("u", "v", "w", "p", "pp"):
        eq(ctx.download(k.upper()), o[k], k)
    eq(ctx.download("FLMASS"), o["flm"], "flmass")
    eq(ctx.download("DPDXI")[: m.numCells], o["dP"][: m.numCells], "dPdxi(pp)")
    for k in ("su", "sv", "sw"):
        eq(ctx.download(k.upper())[: m.numCells], o[k], k)
    ctx.close()


@pytest.mark.parametrize("solver", ["dpcg", "iccg", "bicgstab"])
@pytest.mark.parametrize("n", [5, 16, 33])
def test_csrsolve_poisson(fcp, orc, solver, n):
    """csrsolve on the Poisson known-answer system: iterates bit-identical to the oracle in TREE summation mode,
    counts within +-1 of the gfortran-order (SEQ) oracle."""
    m = M.cavity_mesh(n, bump=0.6)
    c, a, su = cases.poisson_system(m, orc)
    sid = L.SOLVER_ID[solver]
    x = np.zeros(m.numCells)
    rep_o = orc.solve(sid, c.ia, c.ja, a, c.diag, x, su, 2000, 1e-30, 1e-10, orc.SUM_TREE)
    xs = np.zeros(m.numCells)
    rep_s = orc.solve(sid, c.ia, c.ja, a, c.diag, xs, su, 2000, 1e-30, 1e-10, orc.SUM_SEQ)
    ctx = make_ctx(m)
    ctx.upload("A", a); ctx.upload("SU", su); ctx.fill("PP", 0.0)
    rep = ctx.csrsolve(solver, "PP", "SU", 2000, 1e-30, 1e-10)
    assert rep.iters == rep_o.iters
    assert abs(rep.iters - rep_s.iters) <= 1
    eq(ctx.download("PP")[: m.numCells], x, "solution")
    assert (rep.res0, rep.resl, rep.factor, rep.resor) == (rep_o.res0, rep_o.resl, rep_o.factor, rep_o.resor)
    assert L.report_line(rep, "p") == orc.report_line(sid, "p", rep_o)
    r = su - orc.spmv(c.ia, c.ja, a, ctx.download("PP")[: m.numCells].copy())
    assert np.abs(r).sum() <= 1e-9 * np.abs(su).sum()
    ctx.close()


def test_csrsolve_early_return_and_maxiter(fcp, orc):
    """res0 < tol_abs returns immediately leaving fi untouched (linear_solvers.f90:266-270); itr_max is honoured."""
    m = M.cavity_mesh(8)
    c, a, su = cases.poisson_system(m, orc)
    ctx = make_ctx(m)
    ctx.upload("A", a); ctx.upload("SU", su)
    x0 = np.random.default_rng(5).standard_normal(m.numTotal)
    for solver in ("dpcg", "iccg", "bicgstab"):
        ctx.upload("PP", x0)
        rep = ctx.csrsolve(solver, "PP", "SU", 100, 1e30, 1e-10)
        assert rep.iters == 0 and rep.resl == rep.res0
        eq(ctx.download("PP"), x0, "fi untouched")
        assert "No Iterations 0" in L.report_line(rep, "p")
        rep = ctx.csrsolve(solver, "PP", "SU", 3, 1e-30, 1e-14)
        assert rep.iters == 3
        xo = x0[: m.numCells].copy()
        ro = orc.solve(L.SOLVER_ID[solver], c.ia, c.ja, a, c.diag, xo, su, 3, 1e-30, 1e-14, orc.SUM_TREE)
        assert ro.iters == 3
        eq(ctx.download("PP")[: m.numCells], xo, "3 iterations")
    ctx.close()


def test_explicit_csr_golden_5x5(fcp):
    """test/test_linear_solvers_spsolve.f90: the reference's two 5x5 systems through the explicit-CSR signature."""
    with open(os.path.join(cases.GOLDEN, "spsolve_5x5.json")) as fh:
        g = json.load(fh)
    s = L.CsrSolver(g["ioffset"], g["ja"], g["diag"])
    for system, solvers in (("nonsymmetric", ["bicgstab"]), ("spd", ["iccg", "dpcg"])):
        for solver in solvers:
            x = np.zeros(5)
            rep = s.solve(solver, g[system]["a_f32"], x, g[system]["b_f32"], 50, g["tol_abs"], g["tol_rel"])
            assert rep.iters <= 50
            np.testing.assert_allclose(x, g[system]["x"], atol=0.0051 + 2e-3 * np.abs(g[system]["x"]).max())
    s.close()


def test_explicit_csr_matches_oracle(fcp, orc):
    m = M.cavity_mesh(9, distort=0.2)
    c, a, su = cases.poisson_system(m, orc)
    s = L.CsrSolver(c.ia, c.ja, c.diag)
    for solver in ("dpcg", "iccg", "bicgstab"):
        x = np.zeros(m.numCells); xo = np.zeros(m.numCells)
        rep = s.solve(solver, a, x, su, 500, 1e-30, 1e-9)
        ro = orc.solve(L.SOLVER_ID[solver], c.ia, c.ja, a, c.diag, xo, su, 500, 1e-30, 1e-9, orc.SUM_TREE)
        assert rep.iters == ro.iters
        eq(x, xo, solver)
    s.close()


def test_determinism(fcp):
    """SURVEY section 5 (race detection): two runs of the whole step are bitwise equal."""
    m = M.cavity_mesh(20, distort=0.2)
    f = cases.fields(m)
    outs = []
    for _ in range(2):
        ctx = make_ctx(m, f)
        ctx.gradp_and_sources("linear", "P")
        ctx.calcp_simple(solver="dpcg", maxiter=200, tol_abs=1e-30, tol_rel=1e-8, pRefCell=1)
        outs.append([ctx.download(k) for k in ("U", "V", "W", "P", "PP", "FLMASS", "A")])
        ctx.close()
    for a, b in zip(*outs):
        eq(a, b, "run-to-run")


def test_errors(fcp):
    m = M.cavity_mesh(4)
    ctx = make_ctx(m)
    with pytest.raises(L.FcpError):
        ctx.grad(L.GRAD_LSQ, "S0", "G0")          # matrix not created
    with pytest.raises(L.FcpError):
        ctx.csrsolve(9, "PP", "SU", 10, 0.0, 0.1)  # unknown solver
    with pytest.raises(L.FcpError):
        ctx.upload("A", np.zeros(3))
    with pytest.raises(L.FcpError):
        ctx.correct_simple("linear", 0.3, 0)
    ctx.close()
    with pytest.raises(L.FcpError):
        L.CsrSolver([1, 3, 4], [2, 1, 2], [2, 3])   # unsorted row / bad diag

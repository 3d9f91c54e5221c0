"""GPU (B200) parity for the remaining rows of SURVEY.md section 8(a): a9 calcp_piso, a12 QR least-squares gradient,
a13 slope limiters.  Same bar as tests/test_gpu_parity.py: BIT-EXACT against the CPU oracle on the same seeded inputs
(reductions in the oracle's TREE mode), through the C-ABI."""
import numpy as np
import pytest

import cases
from fcb200 import lib as L
from fcb200 import mesh as M
from test_gpu_parity import MESHES, eq, make_ctx

pytestmark = pytest.mark.gpu

LIMITERS = ["Barth-Jespersen", "Venkatakrishnan", "R3", "multidimensional"]


@pytest.fixture(scope="module")
def allmeshes():
    return cases.meshes()


def steep_field(m, seed=5):
    """A field with local extrema and steep fronts, so that every limiter branch (r > 0, r < 0, |delta| < 1e-6) is taken."""
    rng = np.random.default_rng(seed)
    f = m.boundary_values_of(lambda x, y, z: np.tanh(8 * (x - 0.45)) * np.cos(3 * y) + 0.3 * np.sin(5 * z) + 0.0 * x)
    f = f + 0.05 * rng.standard_normal(m.numTotal)
    n = m.numCells
    f[: max(n // 7, 1)] = f[0]          # a flat patch: delta_face ~ 0 branch
    return f


@pytest.mark.parametrize("name", MESHES)
@pytest.mark.parametrize("limiter", LIMITERS)
def test_slope_limiters(fcp, orc, allmeshes, name, limiter):
    """gradients.f90:288-656, global-extrema quirk Q3 reproduced."""
    m = allmeshes[name]
    phi = steep_field(m)
    c = orc.Csr(m)
    g = orc.grad_gauss(m, phi)
    g_lim = orc.slope_limiter(m, c, L.LIMITER_ID[limiter], phi, g.copy())
    ctx = make_ctx(m)
    ctx.upload("S0", phi)
    ctx.grad_opt("gauss", limiter, "S0", "G0")
    eq(ctx.download("G0"), g_lim, f"grad(gauss) + {limiter}")
    assert not np.array_equal(g_lim, g) or m.numCells < 10, "the limiter never acted: the test field is too smooth"
    # the limiter on its own, on a least-squares gradient
    ctx.create_lsq_grad_matrix(L.GRAD_LSQ)
    ctx.grad(L.GRAD_LSQ, "S0", "G1")
    g2 = ctx.download("G1")
    ctx.slope_limiter(limiter, "S0", "G1")
    eq(ctx.download("G1"), orc.slope_limiter(m, c, L.LIMITER_ID[limiter], phi, g2.copy()), f"lsq + {limiter}")
    ctx.close()


@pytest.mark.parametrize("name", MESHES)
def test_grad_lsq_qr(fcp, orc, allmeshes, name):
    """create_matrix_lsq_qr / grad_lsq_qr, gradients.f90:900-1152 (hexahedra: 6 faces per cell)."""
    m = allmeshes[name]
    f = cases.fields(m)
    ctx = make_ctx(m)
    if name.startswith("poly") or name == "hex_many_faces":
        # m = 6 (gradients.f90:924): a cell with more than 6 faces cannot be held -> FCP_EINVAL, not silent corruption
        with pytest.raises(ValueError):
            orc.create_matrix_lsq_qr(m)
        with pytest.raises(L.FcpError):
            ctx.create_lsq_grad_matrix(L.GRAD_LSQ_QR)
        ctx.close()
        return
    D = orc.create_matrix_lsq_qr(m)
    ctx.upload("S0", f["p"])
    ctx.create_lsq_grad_matrix(L.GRAD_LSQ_QR)
    ctx.grad(L.GRAD_LSQ_QR, "S0", "G0")
    eq(ctx.download("G0"), orc.grad_lsq_qr(m, D, f["p"]), "grad_lsq_qr")
    # exact for a linear field
    lin = m.boundary_values_of(lambda x, y, z: 1.0 + x - 2 * y + 3 * z)
    ctx.upload("S0", lin)
    ctx.grad_opt("lsq_qr", "none", "S0", "G0")
    np.testing.assert_allclose(ctx.download("G0")[: m.numCells], np.tile([1.0, -2.0, 3.0], (m.numCells, 1)), rtol=0, atol=5e-11)
    ctx.close()


def piso_inputs(orc, m, f, seed=11):
    """A momentum-like matrix (negative off-diagonals, dominant diagonal), its right-hand sides and reciprocal diagonals."""
    rng = np.random.default_rng(seed)
    c = orc.Csr(m)
    n = m.numCells
    a = -np.abs(rng.standard_normal(c.nnz)) - 0.1
    row = np.repeat(np.arange(n), np.diff(c.ia))
    offsum = np.bincount(row, weights=np.abs(a), minlength=n) - np.abs(a[c.diag - 1])
    a[c.diag - 1] = 1.3 * offsum + 0.5
    ap = np.ones(m.numTotal)
    ap[:n] = 1.0 / a[c.diag - 1]
    scale = (m.vol[:n].mean()) ** (2.0 / 3.0)
    a *= scale
    ap[:n] /= scale
    r = {k: scale * rng.standard_normal(m.numTotal) for k in ("ru", "rv", "rw")}
    return c, a, ap, r


@pytest.mark.parametrize("name", MESHES)
@pytest.mark.parametrize("solver,pscheme,npcor", [("dpcg", "linear", 1), ("iccg", "weighted", 2), ("bicgstab", "central", 2)])
def test_calcp_piso(fcp, orc, allmeshes, name, solver, pscheme, npcor):
    """calcp_piso.f90:81-489: 2 correctors x npcor non-orthogonal passes; every field, the matrix and the counts identical."""
    m = allmeshes[name]
    f = cases.fields(m)
    c, a, ap, r = piso_inputs(orc, m, f)
    sid = L.SOLVER_ID[solver]
    g = {k: v.copy() for k, v in f.items()}
    for k in ("apu", "apv", "apw"):
        g[k] = ap * (1.0 + 0.01 * ("uvw".index(k[-1])))
    dP = np.zeros((m.numTotal, 3))
    orc.gradp_and_sources(m, 0, g["p"], g["apu"], dP)
    flm = np.zeros(m.numFaces)
    Fi = m.numInnerFaces
    for ib in range(m.numBoundaries):
        if m.bctype[ib] == M.BC_INLET:
            pf = m.patch_faces(ib)
            ijb = m.numCells + pf - Fi
            flm[pf] = g["den"][ijb] * (g["u"][ijb] * m.arx[pf] + g["v"][ijb] * m.ary[pf] + g["w"][ijb] * m.arz[pf])
    ctx = make_ctx(m, dict(u=g["u"], v=g["v"], w=g["w"], p=g["p"], pp=g["pp"], den=g["den"], apu=g["apu"], apv=g["apv"], apw=g["apw"]))
    ctx.upload("A", a); ctx.upload("RU", r["ru"]); ctx.upload("RV", r["rv"]); ctx.upload("RW", r["rw"])
    ctx.upload("FLMASS", flm); ctx.upload("DPDXI", dP)
    reps = ctx.calcp_piso(solver=solver, maxiter=40, tol_abs=1e-30, tol_rel=1e-7, urfp=0.9, ncorr=2, npcor=npcor, pscheme=pscheme, flomas=1.0)
    ao = a.copy()
    oreps, su, sv, sw, h = orc.calcp_piso(m, c, sid, 40, 1e-30, 1e-7, orc.SUM_TREE, 2, npcor, L.PSCHEME[pscheme], 0.9, False, 1.0,
                                          r["ru"], r["rv"], r["rw"], g["den"], g["apu"], g["apv"], g["apw"], ao,
                                          g["u"], g["v"], g["w"], g["p"], g["pp"], dP, flm)
    for rg, ro in zip(reps, oreps):
        assert rg.iters == ro.iters, (rg.iters, ro.iters)
        assert rg.res0 == ro.res0 and rg.resl == ro.resl
    eq(ctx.download("H"), h, "h = a (momentum coefficients)")
    eq(ctx.download("A"), ao, "pressure matrix")
    for k in ("u", "v", "w", "p", "pp"):
        eq(ctx.download(k.upper()), g[k], k)
    eq(ctx.download("FLMASS"), flm, "flmass")
    eq(ctx.download("DPDXI")[: m.numCells], dP[: m.numCells], "dPdxi")
    eq(ctx.download("SU")[: m.numCells], su, "su"); eq(ctx.download("SV")[: m.numCells], sv, "sv"); eq(ctx.download("SW")[: m.numCells], sw, "sw")
    ctx.close()


def test_periodic_pair_rules(fcp):
    """A periodic patch must name an 'empty' twin of the same size; a pair must not join cells that already share a face."""
    xs = np.linspace(0, 1, 5)
    bad = M.hex_mesh(xs, xs, xs, dict(left="empty", right="periodic"))
    bad.startFaceTwin = None
    with pytest.raises(L.FcpError):
        make_ctx(bad)
    two = M.hex_mesh(np.linspace(0, 1, 3), xs, xs, dict(left="empty", right="periodic"))    # 2 cells across: the pair duplicates an inner face
    with pytest.raises(L.FcpError):
        make_ctx(two)


# ---- row f1: the momentum predictor -----------------------------------------------------------------------------------------
def uvw_inputs(m, seed=21, transient=0):
    rng = np.random.default_rng(seed)
    f = cases.fields(m)
    n, nT, Fi = m.numCells, m.numTotal, m.numInnerFaces
    g = {k: f[k].copy() for k in ("u", "v", "w", "p", "den", "apu")}
    g["vis"] = 0.01 * m.boundary_values_of(lambda x, y, z: 1.0 + 0.5 * np.sin(3 * x) * np.cos(2 * y) + 0.2 * z)
    g["visw"] = 0.01 * (1.0 + rng.random(m.numBoundaryFaces))             # some above, some below `viscos`
    # a mass-flux field with both signs (upwinding takes both branches)
    own, nb = m.owner[:Fi].astype(np.int64) - 1, m.neighbour.astype(np.int64) - 1
    uf = 0.5 * (g["u"][own] + g["u"][nb]); vf = 0.5 * (g["v"][own] + g["v"][nb]); wf = 0.5 * (g["w"][own] + g["w"][nb])
    flm = np.zeros(m.numFaces)
    flm[:Fi] = uf * m.arx[:Fi] + vf * m.ary[:Fi] + wf * m.arz[:Fi]
    ib = n + np.arange(m.numBoundaryFaces)
    flm[Fi:] = g["den"][ib] * (g["u"][ib] * m.arx[Fi:] + g["v"][ib] * m.ary[Fi:] + g["w"][ib] * m.arz[Fi:])
    for ipatch in range(m.numBoundaries):
        if m.bctype[ipatch] in (M.BC_WALL, M.BC_SYMMETRY, M.BC_EMPTY):
            flm[m.patch_faces(ipatch)] = 0.0
    g["flmass"] = flm
    for lvl in range(transient):
        for c in "uvw":
            g[c + "o" * (lvl + 1)] = g[c] * (1.0 - 0.05 * (lvl + 1)) + 1e-3 * rng.standard_normal(nT)
    return g


@pytest.mark.parametrize("name", MESHES)
@pytest.mark.parametrize("cscheme,grad,limiter,tscheme,solver,pscheme", [
    ("cds", "gauss", "none", "steady", "bicgstab", "linear"),
    ("muscl", "gauss", "Venkatakrishnan", "bdf", "bicgstab", "weighted"),
    ("linearUpwind", "lsq", "Barth-Jespersen", "bdf2", "bicgstab", "linear"),
    ("central", "wlsq", "none", "bdf3", "bicgstab", "central"),
])
def test_calcuvw(fcp, orc, allmeshes, name, cscheme, grad, limiter, tscheme, solver, pscheme):
    """calcuvw (velocity.f90:50-750): the matrix, sources, reciprocal diagonals, gradients and the three BiCGStab-ILU(0) solves
    bit-identical to the oracle (reductions in TREE mode), iteration counts identical."""
    m = allmeshes[name]
    nlev = L.TSCHEME[tscheme]
    g = uvw_inputs(m, transient=nlev)
    c = orc.Csr(m)
    a0 = np.random.default_rng(4).standard_normal(c.nnz)          # stale matrix: its diagonal enters the first row sum (:606)
    prm = orc.OrcUvwParams()
    prm.solver, prm.maxiter, prm.tol_abs, prm.tol_rel = L.SOLVER_ID[solver], 6, 1e-30, 1e-3
    prm.urf[0], prm.urf[1], prm.urf[2] = 0.8, 0.7, 0.75
    prm.gds, prm.cscheme = 0.9, L.CSCHEME_ID[cscheme]
    prm.grad_method, prm.limiter, prm.pscheme = L.GRAD_ID[grad], L.LIMITER_ID[limiter], L.PSCHEME[pscheme]
    prm.tscheme, prm.timestep, prm.piso, prm.const_mflux, prm.gradPcmf, prm.viscos = nlev, 0.01, 1, 1, 0.3, 0.015
    prm.sum_mode = orc.SUM_TREE
    ctx = make_ctx(m, {k: g[k] for k in ("u", "v", "w", "p", "den", "apu", "vis")})
    visw_field = np.zeros(m.numTotal); visw_field[m.numCells:] = g["visw"]
    ctx.upload("VISW", visw_field); ctx.upload("FLMASS", g["flmass"]); ctx.upload("A", a0)
    for k in g:
        if k[0] in "uvw" and k.endswith("o"):
            ctx.upload(k.upper(), g[k])
    reps = ctx.calcuvw(solver=solver, maxiter=6, tol_abs=1e-30, tol_rel=1e-3, urf=(0.8, 0.7, 0.75), gds=0.9, cscheme=cscheme, grad_method=grad,
                       limiter=limiter, pscheme=pscheme, tscheme=tscheme, timestep=0.01, piso=True, const_mflux=True, gradPcmf=0.3, viscos=0.015)
    a = a0.copy()
    o = orc.calcuvw(m, c, prm, g, a)
    n = m.numCells
    eq(ctx.download("DUDXI")[:n], o["dUdxi"][:n], "dUdxi"); eq(ctx.download("DWDXI")[:n], o["dWdxi"][:n], "dWdxi")
    eq(ctx.download("RU")[:n], o["rU"], "rU"); eq(ctx.download("RV")[:n], o["rV"], "rV"); eq(ctx.download("RW")[:n], o["rW"], "rW")
    eq(ctx.download("SPU")[:n], o["spu"], "spu"); eq(ctx.download("SP")[:n], o["sp"], "sp")
    eq(ctx.download("APU")[:n], o["apu"][:n], "apu"); eq(ctx.download("APV")[:n], o["apv"][:n], "apv"); eq(ctx.download("APW")[:n], o["apw"][:n], "apw")
    eq(ctx.download("A"), a, "a (W-equation matrix)")
    for r, ro in zip(reps, o["reps"]):
        assert r.iters == ro.iters and r.res0 == ro.res0 and r.resl == ro.resl, ((r.iters, r.res0, r.resl), (ro.iters, ro.res0, ro.resl))
    for k in ("u", "v", "w", "p"):
        eq(ctx.download(k.upper()), g[k], k)
    eq(ctx.download("SU")[:n], o["su"], "su (W right-hand side)")
    ctx.close()


@pytest.mark.parametrize("cscheme", L.CSCHEMES)
def test_calcuvw_all_convection_schemes(fcp, orc, allmeshes, cscheme):
    """Every cSchemeU string of interpolation.f90:28-113 / :596-640, single-precision literals (2./3., 1./3., 1e-30) included."""
    m = allmeshes["channel_inout"]
    g = uvw_inputs(m)
    c = orc.Csr(m)
    prm = orc.OrcUvwParams()
    prm.solver, prm.maxiter, prm.tol_abs, prm.tol_rel = L.SOLVER_BICGSTAB, 3, 1e-30, 1e-2
    prm.urf[0] = prm.urf[1] = prm.urf[2] = 0.8
    prm.gds, prm.cscheme, prm.viscos, prm.sum_mode = 1.0, L.CSCHEME_ID[cscheme], 0.01, orc.SUM_TREE
    ctx = make_ctx(m, {k: g[k] for k in ("u", "v", "w", "p", "den", "apu", "vis")})
    visw_field = np.zeros(m.numTotal); visw_field[m.numCells:] = g["visw"]
    ctx.upload("VISW", visw_field); ctx.upload("FLMASS", g["flmass"])
    ctx.calcuvw(solver="bicgstab", maxiter=3, tol_abs=1e-30, tol_rel=1e-2, urf=(0.8, 0.8, 0.8), gds=1.0, cscheme=cscheme, viscos=0.01)
    a = np.zeros(c.nnz)
    o = orc.calcuvw(m, c, prm, g, a)
    for k in ("u", "v", "w"):
        eq(ctx.download(k.upper()), g[k], f"{k} with {cscheme}")
    ctx.close()


# ---- row f3: periodic channels -------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["channel_periodic", "hex10_distorted"])
def test_constant_mass_flow_forcing(fcp, orc, allmeshes, name):
    """constant_mass_flow_forcing.f90: bit-identical to the oracle with the sums in TREE mode, to rounding in the reference's cell-by-cell order."""
    m = allmeshes[name]
    f = cases.fields(m)
    ctx = make_ctx(m, dict(u=f["u"], apu=f["apu"]))
    g_new, ustar = ctx.constant_mass_flow_forcing(0.1335, 0.25)
    u = f["u"].copy()
    gplus, ustar_o = orc.constant_mass_flow_forcing(m, 0.1335, f["apu"], u, orc.SUM_TREE)
    assert ustar == ustar_o and g_new == 0.25 + gplus
    eq(ctx.download("U"), u, "u after the forcing correction")
    u2 = f["u"].copy()
    gseq, ustar_s = orc.constant_mass_flow_forcing(m, 0.1335, f["apu"], u2, orc.SUM_SEQ)
    assert abs(gseq - gplus) <= 1e-12 * abs(gplus) and abs(ustar_s - ustar) <= 1e-13 * abs(ustar)
    # the corrected field has the requested bulk velocity
    n = m.numCells
    assert abs((m.vol[:n] * u[:n]).sum() / m.vol[:n].sum() - 0.1335) < 1e-12
    ctx.close()


@pytest.mark.parametrize("name", ["channel_periodic", "duct_periodic_x", "duct_periodic_first", "channel_inout", "channel_pressure"])
def test_update_boundary(fcp, orc, allmeshes, name):
    m = allmeshes[name]
    phi = np.random.default_rng(8).standard_normal(m.numTotal)
    ctx = make_ctx(m, dict(s0=phi))
    ctx.update_boundary("S0")
    ref = phi.copy()
    orc.update_boundary(m, ref)
    eq(ctx.download("S0"), ref, "updateBoundary")
    ctx.close()


@pytest.mark.parametrize("name", ["hex10_distorted", "hex12_graded", "channel_inout", "poly_10faces", "channel_pressure"])
@pytest.mark.parametrize("method", ["gauss", "lsq"])
def test_assemble_pcorr_with_the_mpi_trees_facefluxmass(fcp, orc, allmeshes, name, method):
    """Quirk Q10 as a switch (`fcp_set_flux_variant(1, grad_method)`): the inner faces of the SIMPLE p' assembly take `facefluxmass` of the MPI tree
    (src-par/calcp_simple.f90:40-79, src-par/faceflux_mass.f90:28-180: face_value_central velocities from grad(U,V,W), per-component (Vol/Ap)_f, the
    P'/E' pressure correction incl. its sign quirk Q26) instead of the serial tree's `facefluxmass2`.  Matrix, sources and fluxes bit-identical to the
    oracle's restatement; the matrix stays symmetric with zero row sums; on the orthogonal graded mesh the two variants' COEFFICIENTS agree to
    rounding (the Rhie-Chow fluxes differ by construction)."""
    import cases
    from fcb200 import lib as L
    m = allmeshes[name]
    f = cases.fields(m)
    n = m.numCells
    gm = dict(gauss=L.GRAD_GAUSS, lsq=L.GRAD_LSQ)[method]
    c = orc.Csr(m)
    g = {k: v.copy() for k, v in f.items()}
    dP = np.zeros((m.numTotal, 3))
    orc.gradp_and_sources(m, 0, g["p"], g["apu"], dP)
    if method == "gauss":
        gU, gV, gW = (orc.grad_gauss(m, g[k]) for k in "uvw")
    else:
        D = orc.create_matrix_lsq(m, False)
        gU, gV, gW = (orc.grad_lsq(m, False, D, g[k]) for k in "uvw")
    a = np.zeros(c.nnz); su = np.zeros(n); flm = np.zeros(m.numFaces)
    Fi = m.numInnerFaces
    for ib in range(m.numBoundaries):
        if m.bctype[ib] == M.BC_INLET:
            pf = m.patch_faces(ib)
            ijb = n + pf - Fi
            flm[pf] = g["den"][ijb] * (g["u"][ijb] * m.arx[pf] + g["v"][ijb] * m.ary[pf] + g["w"][ijb] * m.arz[pf])
    flm0 = flm.copy()
    orc.assemble_pcorr_mpi_into(m, c, g["den"], g["u"], g["v"], g["w"], g["p"], g["pp"], dP, g["apu"], g["apv"], g["apw"], gU, gV, gW, a, su, flm, flomas=1.0)
    ctx = L.Context(m)
    for k, v in f.items():
        ctx.upload(k.upper(), v)
    ctx.upload("FLMASS", flm0)
    if method == "lsq":
        ctx.create_lsq_grad_matrix(gm)
    ctx.gradp_and_sources("linear", "P")
    ctx.set_flux_variant(1, gm)
    ctx.assemble_pcorr_simple(False, 1.0)
    got_a, got_su, got_fl = ctx.download("A"), ctx.download("SU")[:n], ctx.download("FLMASS")
    assert np.array_equal(got_a, a), np.abs(got_a - a).max()
    assert np.array_equal(got_su, su) and np.array_equal(got_fl, flm)
    assert np.array_equal(a[c.icell_jcell[:Fi] - 1], a[c.jcell_icell[:Fi] - 1])
    ctx.set_flux_variant(0)
    ctx.upload("FLMASS", flm0)
    for k in ("u", "v", "w", "pp"):
        ctx.upload(k.upper(), f[k])
    ctx.assemble_pcorr_simple(False, 1.0)
    a2 = ctx.download("A")
    assert not np.array_equal(ctx.download("FLMASS")[:Fi], flm[:Fi])
    if name == "hex12_graded":
        assert np.abs(a2 - a).max() <= 1e-12 * np.abs(a).max()
    with pytest.raises(L.FcpError):
        ctx.set_flux_variant(2)
    ctx.close()

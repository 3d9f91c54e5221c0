"""GPU (B200): row f4 -- the scalar transport template (fcp_calcsc: scalar_fluxes.f90 inside the calcsc assembly of k_epsilon_rlzb.f90),
calc_strain_and_vorticity and modify_mu_eff against the oracle.  Bit-exact wherever only + - * / sqrt are involved (the generic scalar and
the k equation); the epsilon equation (k**1.5 in the wall cells) and modify_mu_eff (acos, cos, log) are compared to 1e-12 relative because
the device's libm and the host's need not round those functions identically."""
import numpy as np
import pytest

import cases
from fcb200 import lib as L
from fcb200 import mesh as M
from test_gpu_parity import eq, make_ctx
from test_gpu_rows2 import uvw_inputs

pytestmark = pytest.mark.gpu
SC_MESHES = ["hex10_distorted", "channel_inout", "channel_pressure", "poly_10faces", "hex_many_faces", "channel_periodic", "duct_periodic_first", "tiny3"]


@pytest.fixture(scope="module")
def allmeshes():
    return cases.meshes()


def close(a, b, what, rtol=1e-12):
    a = np.asarray(a); b = np.asarray(b)
    err = np.abs(a - b).max() / (np.abs(b).max() + 1e-300)
    assert err <= rtol, f"{what}: relative difference {err:.3e}"


def scalar_inputs(m, orc, seed=5):
    """k, epsilon, effective viscosity, wall distance, strain ... for one mesh (all positive where the model divides by them)."""
    rng = np.random.default_rng(seed)
    g = uvw_inputs(m)
    n, nT, B = m.numCells, m.numTotal, m.numBoundaryFaces
    g["te"] = 0.02 * m.boundary_values_of(lambda x, y, z: 1.0 + 0.5 * np.sin(2 * x) * np.cos(3 * y) + 0.2 * z) + 1e-4 * rng.random(nT)
    g["ed"] = 0.05 * m.boundary_values_of(lambda x, y, z: 1.0 + 0.4 * np.cos(x + 2 * y) + 0.1 * np.sin(3 * z)) + 1e-4 * rng.random(nT)
    g["vis"] = 0.01 + 0.03 * m.boundary_values_of(lambda x, y, z: 1.0 + 0.5 * np.sin(3 * x) * np.cos(2 * y) + 0.2 * z)
    dnw, srdw, dns, srds = orc.wall_geometry(m)
    g["dnw"] = np.full(B, 0.05)
    iw = 0
    for ib in range(m.numBoundaries):
        if m.bctype[ib] == M.BC_WALL:
            pf = m.patch_faces(ib) - m.numInnerFaces
            g["dnw"][pf] = dnw[iw: iw + pf.size]
            iw += pf.size
    gU = orc.grad_gauss(m, g["u"]); gV = orc.grad_gauss(m, g["v"]); gW = orc.grad_gauss(m, g["w"])
    g["gU"], g["gV"], g["gW"] = gU, gV, gW
    g["magStrain"], g["vorticity"] = orc.calc_strain_and_vorticity(m, gU, gV, gW)
    g["phio"] = g["te"] * 0.97 + 1e-5 * rng.random(nT)
    g["phioo"] = g["te"] * 0.94 + 1e-5 * rng.random(nT)
    return g


def bslot(m, arr):
    """per-boundary-face array -> numTotal field with the values in the boundary slots"""
    out = np.zeros(m.numTotal)
    out[m.numCells:] = arr
    return out


def upload_scalar_state(ctx, m, g):
    for k in ("u", "v", "w", "den", "vis", "te", "ed", "phio", "phioo"):
        ctx.upload(k.upper(), g[k])
    ctx.upload("FLMASS", g["flmass"])
    ctx.upload("VISW", bslot(m, g["visw"])); ctx.upload("DNW", bslot(m, g["dnw"]))
    ms = np.zeros(m.numTotal); ms[: m.numCells] = g["magStrain"]
    ctx.upload("MAGSTRAIN", ms)


def oracle_params(orc, kind, solver, cscheme, grad, limiter, tscheme):
    prm = orc.OrcScalarParams()
    prm.kind, prm.solver, prm.maxiter, prm.cscheme = kind, L.SOLVER_ID[solver], 8, L.CSCHEME_ID[cscheme]
    prm.grad_method, prm.limiter, prm.tscheme, prm.sum_mode = L.GRAD_ID[grad], L.LIMITER_ID[limiter], L.TSCHEME[tscheme], orc.SUM_TREE
    prm.tol_abs, prm.tol_rel, prm.urf, prm.gds, prm.timestep, prm.prtr, prm.viscos, prm.densit = 1e-30, 1e-4, 0.7, 0.8, 0.02, 1.0 / 1.2, 0.01, 1.0
    return prm


@pytest.mark.parametrize("name", SC_MESHES)
def test_strain_and_vorticity(fcp, orc, allmeshes, name):
    m = allmeshes[name]
    g = scalar_inputs(m, orc)
    ctx = make_ctx(m)
    ctx.upload("DUDXI", g["gU"]); ctx.upload("DVDXI", g["gV"]); ctx.upload("DWDXI", g["gW"])
    ctx.calc_strain_and_vorticity()
    eq(ctx.download("MAGSTRAIN")[: m.numCells], g["magStrain"], "magStrain")
    eq(ctx.download("VORTICITY")[: m.numCells], g["vorticity"], "vorticity")
    ctx.close()


@pytest.mark.parametrize("name", SC_MESHES)
@pytest.mark.parametrize("cscheme,grad,limiter,tscheme", [("cds", "gauss", "none", "steady"), ("muscl", "gauss", "Venkatakrishnan", "bdf"),
                                                          ("linearUpwind", "lsq", "none", "bdf2")])
def test_calcsc_generic(fcp, orc, allmeshes, name, cscheme, grad, limiter, tscheme):
    """A passive scalar with caller-supplied volume sources: matrix, sources, gradient, solution and the iteration count bit-identical."""
    m = allmeshes[name]
    g = scalar_inputs(m, orc)
    n = m.numCells
    rng = np.random.default_rng(11)
    phi = m.boundary_values_of(lambda x, y, z: 1.0 + 0.3 * np.sin(2 * x + y) + 0.1 * z)
    suv = np.zeros(m.numTotal); suv[:n] = 0.1 * m.vol[:n] * rng.random(n)
    spv = np.zeros(m.numTotal); spv[:n] = 0.5 * m.vol[:n] * rng.random(n)
    ctx = make_ctx(m)
    upload_scalar_state(ctx, m, g)
    ctx.upload("S0", phi); ctx.upload("S2", suv); ctx.upload("S3", spv)
    rep, lo, hi = ctx.calcsc("S0", kind="generic", solver="bicgstab", maxiter=8, tol_abs=1e-30, tol_rel=1e-4, urf=0.7, gds=0.8, cscheme=cscheme,
                             grad_method=grad, limiter=limiter, tscheme=tscheme, timestep=0.02, prtr=1.0 / 1.2, viscos=0.01, densit=1.0)
    c = orc.Csr(m)
    prm = oracle_params(orc, orc.SC_GENERIC, "bicgstab", cscheme, grad, limiter, tscheme)
    f = dict(g, phi=phi.copy(), su_vol=suv[:n].copy(), sp_vol=spv[:n].copy())
    o = orc.calcsc(m, c, prm, f)
    assert (rep.iters, rep.res0, rep.resl) == (o["rep"].iters, o["rep"].res0, o["rep"].resl)
    eq(ctx.download("G0")[:n], o["grad"][:n], "grad(phi)")
    eq(ctx.download("A"), o["a"], "matrix")
    eq(ctx.download("SP")[:n], o["sp"], "sp")
    eq(ctx.download("SU")[:n], o["su"], "su")
    eq(ctx.download("S0"), f["phi"], "phi")
    assert (lo, hi) == (o["fimin"], o["fimax"])
    ctx.close()


@pytest.mark.parametrize("name", SC_MESHES)
@pytest.mark.parametrize("cscheme,tscheme,solver", [("cds", "steady", "bicgstab"), ("muscl", "bdf2", "bicgstab")])
def test_calcsc_tke(fcp, orc, allmeshes, name, cscheme, tscheme, solver):
    """calcsc_tke (k_epsilon_rlzb.f90:52-445): production, the wall-function production in wall cells, tau, the solve: bit-identical."""
    m = allmeshes[name]
    g = scalar_inputs(m, orc)
    n = m.numCells
    ctx = make_ctx(m)
    upload_scalar_state(ctx, m, g)
    rep, lo, hi = ctx.calcsc("TE", kind="tke_rlzb", solver=solver, maxiter=8, tol_abs=1e-30, tol_rel=1e-4, urf=0.7, gds=0.8, cscheme=cscheme,
                             tscheme=tscheme, timestep=0.02, prtr=1.0, viscos=0.01, densit=1.0)
    c = orc.Csr(m)
    prm = oracle_params(orc, orc.SC_TKE_RLZB, solver, cscheme, "gauss", "none", tscheme)
    prm.prtr = 1.0
    f = {k: (v.copy() if isinstance(v, np.ndarray) else v) for k, v in g.items()}
    o = orc.calcsc(m, c, prm, f)
    assert (rep.iters, rep.res0, rep.resl) == (o["rep"].iters, o["rep"].res0, o["rep"].resl)
    eq(ctx.download("A"), o["a"], "matrix")
    eq(ctx.download("SP")[:n], o["sp"], "sp"); eq(ctx.download("SU")[:n], o["su"], "su")
    eq(ctx.download("GEN")[:n], o["gen"], "gen")
    eq(ctx.download("TAU")[n:], o["tau"][: m.numBoundaryFaces], "tau")
    eq(ctx.download("TE"), f["te"], "te")
    assert (lo, hi) == (o["fimin"], o["fimax"])
    ctx.close()


@pytest.mark.parametrize("name", SC_MESHES)
@pytest.mark.parametrize("cscheme,tscheme", [("cds", "steady"), ("vanleer", "bdf")])
def test_calcsc_epsilon(fcp, orc, allmeshes, name, cscheme, tscheme):
    """calcsc_epsilon (:447-790): realizable c1, cleared rows and the imposed epsilon in wall cells.  k**1.5 goes through pow(): 1e-12."""
    m = allmeshes[name]
    g = scalar_inputs(m, orc)
    g["phio"] = g["ed"] * 0.97; g["phioo"] = g["ed"] * 0.95
    n = m.numCells
    ctx = make_ctx(m)
    upload_scalar_state(ctx, m, g)
    rep, lo, hi = ctx.calcsc("ED", kind="eps_rlzb", solver="bicgstab", maxiter=8, tol_abs=1e-30, tol_rel=1e-4, urf=0.7, gds=0.8, cscheme=cscheme,
                             tscheme=tscheme, timestep=0.02, prtr=1.0 / 1.2, viscos=0.01, densit=1.0)
    c = orc.Csr(m)
    prm = oracle_params(orc, orc.SC_EPS_RLZB, "bicgstab", cscheme, "gauss", "none", tscheme)
    f = {k: (v.copy() if isinstance(v, np.ndarray) else v) for k, v in g.items()}
    o = orc.calcsc(m, c, prm, f)
    # on a mesh whose every cell is a wall cell (tiny3) the system is the identity with the imposed value on both sides: the initial residual is
    # zero or one rounding error of pow() -- then the count is 0 or 1 depending on the libm's last bit (found with FCP_EMU_LIBM_ULP=3)
    degenerate = max(rep.res0, o["rep"].res0) <= 1e-13 * np.abs(f["ed"][:n]).sum()
    assert rep.iters == o["rep"].iters or (degenerate and abs(rep.iters - o["rep"].iters) <= 1)
    close(ctx.download("A"), o["a"], "matrix")
    close(ctx.download("SP")[:n], o["sp"], "sp"); close(ctx.download("SU")[:n], o["su"], "su")
    close(ctx.download("ED"), f["ed"], "ed", 1e-10)
    # wall cells hold the imposed value and an identity row
    wall_cells = np.unique(np.concatenate([m.owner[m.patch_faces(ib)] - 1 for ib in range(m.numBoundaries) if m.bctype[ib] == M.BC_WALL] + [np.zeros(0, int)])).astype(int)
    if wall_cells.size and m.numPeriodic == 0:      # (a periodic patch listed after the wall patch re-fills its two entries, as in the reference)
        rows = np.repeat(np.arange(n), np.diff(c.ia))
        a = ctx.download("A")
        offd = np.ones(c.nnz, bool); offd[c.diag - 1] = False
        assert np.all(a[offd & np.isin(rows, wall_cells)] == 0.0)
    ctx.close()


@pytest.mark.parametrize("name", SC_MESHES)
def test_modify_mu_eff_rlzb(fcp, orc, allmeshes, name):
    m = allmeshes[name]
    g = scalar_inputs(m, orc)
    n = m.numCells
    ctx = make_ctx(m)
    upload_scalar_state(ctx, m, g)
    ctx.upload("DUDXI", g["gU"]); ctx.upload("DVDXI", g["gV"]); ctx.upload("DWDXI", g["gW"])
    ctx.modify_mu_eff_k_epsilon_rlzb(0.6, 0.01)
    vis, visw = g["vis"].copy(), g["visw"].copy()
    ypl, tau = orc.modify_mu_eff_rlzb(m, 0.6, 0.01, g["gU"], g["gV"], g["gW"], g["te"], g["ed"], g["den"], g["u"], g["v"], g["w"], g["dnw"], vis, visw)
    close(ctx.download("VIS"), vis, "vis")
    wall = np.zeros(m.numBoundaryFaces, bool)
    for ib in range(m.numBoundaries):
        if m.bctype[ib] == M.BC_WALL:
            wall[m.patch_faces(ib) - m.numInnerFaces] = True
    if wall.any():
        close(ctx.download("VISW")[n:][wall], visw[wall], "visw")
        close(ctx.download("YPL")[n:][wall], ypl[: m.numBoundaryFaces][wall], "ypl")
        close(ctx.download("TAU")[n:][wall], tau[: m.numBoundaryFaces][wall], "tau")
    ctx.close()


def test_calcsc_argument_checks(fcp, allmeshes):
    ctx = make_ctx(allmeshes["tiny3"])
    with pytest.raises(L.FcpError):
        ctx.calcsc("ED", kind="tke_rlzb")
    with pytest.raises(L.FcpError):
        ctx.calcsc("S0", kind="generic", tscheme="bdf3")
    with pytest.raises(L.FcpError):
        ctx.calcsc("DPDXI", kind="generic")
    ctx.close()


# ---- LES sub-grid viscosity -----------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", SC_MESHES)
def test_grad_gauss_fvx(fcp, orc, allmeshes, name):
    """fvxGradient.f90:1549-1662: the two-pass Gauss gradient of the tensor-field layer, bit-identical."""
    m = allmeshes[name]
    g = scalar_inputs(m, orc)
    ctx = make_ctx(m, dict(u=g["u"]))
    ctx.grad_gauss_fvx("U", "DUDXI")
    gx, gy, gz = orc.grad_gauss_fvx(m, g["u"])
    got = ctx.download("DUDXI")[: m.numCells]
    eq(got[:, 0], gx, "dudx"); eq(got[:, 1], gy, "dudy"); eq(got[:, 2], gz, "dudz")
    ctx.close()


@pytest.mark.parametrize("name", SC_MESHES)
@pytest.mark.parametrize("nigrad", [1, 2, 3, 4])
def test_grad_gauss_iter(fcp, orc, allmeshes, name, nigrad):
    """The MPI tree's grad_gauss (src-par/gradients.f90:1547-1664): `nigrad` passes of gradco, bit-identical to the oracle for every pass count (odd
    and even: the passes alternate between the result field and the scratch field); nigrad = 2 is the fvx gradient; nigrad = 1 is the plain Gauss
    gradient up to the rounding of gradco's P fxp + N fxn against grad_gauss' P + (N - P) lambda."""
    from fcb200 import lib as L
    m = allmeshes[name]
    g = scalar_inputs(m, orc)
    ctx = make_ctx(m, dict(u=g["u"]))
    ctx.grad_gauss_iter("U", "DUDXI", nigrad)
    gx, gy, gz = orc.grad_gauss_iter(m, g["u"], nigrad)
    got = ctx.download("DUDXI")[: m.numCells]
    eq(got[:, 0], gx, "dudx"); eq(got[:, 1], gy, "dudy"); eq(got[:, 2], gz, "dudz")
    if nigrad == 2:
        fx, fy, fz = orc.grad_gauss_fvx(m, g["u"])
        assert np.array_equal(fx, gx) and np.array_equal(fy, gy) and np.array_equal(fz, gz)
    if nigrad == 1:
        ctx.grad(L.GRAD_GAUSS, "U", "G0")
        plain = ctx.download("G0")[: m.numCells]
        assert np.abs(plain - got).max() <= 1e-12 * np.abs(plain).max()
    with pytest.raises(L.FcpError):
        ctx.grad_gauss_iter("U", "DUDXI", 0)
    ctx.close()


@pytest.mark.parametrize("name", SC_MESHES)
@pytest.mark.parametrize("model", ["wale", "vreman"])
def test_modify_viscosity_sgs(fcp, orc, allmeshes, name, model):
    """wale_sgs.f90 / vremanSGS.f90 through the tensorFields algebra (pow() involved: 1e-12)."""
    m = allmeshes[name]
    g = scalar_inputs(m, orc)
    n = m.numCells
    ctx = make_ctx(m, dict(u=g["u"], v=g["v"], w=g["w"], den=g["den"], vis=g["vis"]))
    ctx.upload("VISW", bslot(m, g["visw"]))
    ctx.modify_viscosity_sgs(model, 0.7, 0.01)
    vis, visw = g["vis"].copy(), g["visw"].copy()
    orc.modify_viscosity_sgs(m, {"wale": orc.SGS_WALE, "vreman": orc.SGS_VREMAN}[model], 0.7, 0.01, g["u"], g["v"], g["w"], g["den"], vis, visw)
    close(ctx.download("VIS"), vis, f"vis ({model})")
    wall = np.zeros(m.numBoundaryFaces, bool)
    for ib in range(m.numBoundaries):
        if m.bctype[ib] == M.BC_WALL:
            wall[m.patch_faces(ib) - m.numInnerFaces] = True
    if wall.any():
        eq(ctx.download("VISW")[n:][wall], visw[wall], "visw")
    assert np.all(vis[:n] >= 0.0)
    ctx.close()


# ---- k-omega SST -------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", SC_MESHES)
@pytest.mark.parametrize("lowre,tscheme,cscheme", [(False, "steady", "cds"), (True, "bdf", "muscl")])
def test_k_omega_sst_pair(fcp, orc, allmeshes, name, lowre, tscheme, cscheme):
    """modify_viscosity_k_omega_sst (k_omega_SST.f90:62-88): calcsc(k), calcsc(omega), modify_mu_eff in sequence, twice (the second k call
    reads the F1 of the first omega call).  The k equation is libm-free and bit-identical on the first pass; F1 and F2 go through tanh, the
    wall functions through log, so everything downstream is compared at 1e-10."""
    m = allmeshes[name]
    g = scalar_inputs(m, orc)
    n, nT = m.numCells, m.numTotal
    g["ed"] = 40.0 * g["ed"]                                   # omega-like magnitudes
    g["walldist"] = np.zeros(nT); g["walldist"][:n] = 0.02 + np.minimum(np.abs(m.yc[:n] - m.yc[:n].min()), np.abs(m.yc[:n].max() - m.yc[:n]))
    g["fsst"] = np.zeros(nT)
    g["lowre"] = int(lowre)
    ctx = make_ctx(m)
    upload_scalar_state(ctx, m, g)
    ctx.upload("WALLDIST", g["walldist"])
    c = orc.Csr(m)
    f = {k: (v.copy() if isinstance(v, np.ndarray) else v) for k, v in g.items()}
    f["gen"] = np.zeros(n)
    for rep_no in range(2):
        common = dict(solver="bicgstab", maxiter=6, tol_abs=1e-30, tol_rel=1e-30, urf=0.7, gds=0.8, cscheme=cscheme, tscheme=tscheme, timestep=0.02,
                      viscos=0.01, densit=1.0, lowre=lowre)
        ctx.upload("PHIO", f["te"] * 0.97); ctx.upload("PHIOO", f["te"] * 0.95)
        rk, lok, hik = ctx.calcsc("TE", kind="tke_sst", **common)
        gen_dev = ctx.download("GEN")[:n]
        a_k = ctx.download("A")
        ctx.upload("PHIO", f["ed"] * 0.97); ctx.upload("PHIOO", f["ed"] * 0.95)
        ro, loo, hio = ctx.calcsc("ED", kind="omega_sst", **common)
        ctx.modify_mu_eff_k_omega_sst(0.8, 0.01, 1.0, lowre)
        # ---- oracle
        prm = oracle_params(orc, orc.SC_TKE_SST, "bicgstab", cscheme, "gauss", "none", tscheme)
        prm.maxiter, prm.tol_rel = 6, 1e-30
        f["phio"], f["phioo"] = f["te"] * 0.97, f["te"] * 0.95
        ok = orc.calcsc(m, c, prm, f)
        f["gen"] = ok["gen"]; f["dTEdxi"] = ok["grad"]
        prm.kind = orc.SC_OMEGA_SST
        f["phio"], f["phioo"] = f["ed"] * 0.97, f["ed"] * 0.95
        oo = orc.calcsc(m, c, prm, f)
        ypl, tau = orc.modify_mu_eff_sst(m, 0.8, 0.01, 1.0, lowre, f["magStrain"], f["walldist"], f["te"], f["ed"], f["den"], f["u"], f["v"], f["w"],
                                         f["dnw"], f["vis"], f["visw"])
        if rep_no == 0:
            eq(a_k, ok["a"], "k matrix (first pass)")
            eq(gen_dev, ok["gen"], "gen (first pass)")
            assert (rk.iters, rk.res0, rk.resl) == (ok["rep"].iters, ok["rep"].res0, ok["rep"].resl)
        assert rk.iters == ok["rep"].iters and ro.iters == oo["rep"].iters
        close(ctx.download("FSST")[:n], f["fsst"][:n], f"pass {rep_no}: F1", 1e-10)
        close(ctx.download("TE"), f["te"], f"pass {rep_no}: k", 1e-10)
        close(ctx.download("ED"), f["ed"], f"pass {rep_no}: omega", 1e-9)
        close(ctx.download("VIS"), f["vis"], f"pass {rep_no}: vis", 1e-9)
        close(np.array([lok, hik, loo, hio]), np.array([ok["fimin"], ok["fimax"], oo["fimin"], oo["fimax"]]), "extrema", 1e-9)
    wall = np.zeros(m.numBoundaryFaces, bool)
    for ib in range(m.numBoundaries):
        if m.bctype[ib] == M.BC_WALL:
            wall[m.patch_faces(ib) - m.numInnerFaces] = True
    if wall.any():
        close(ctx.download("VISW")[n:][wall], f["visw"][wall], "visw", 1e-9)
        close(ctx.download("YPL")[n:][wall], ypl[: m.numBoundaryFaces][wall], "ypl", 1e-9)
        close(ctx.download("TAU")[n:][wall], tau[: m.numBoundaryFaces][wall], "tau", 1e-9)
    ctx.close()


@pytest.mark.parametrize("name", ["hex10_distorted", "channel_inout", "poly_10faces", "channel_periodic"])
def test_wall_distance(fcp, orc, allmeshes, name):
    """src/mesh/wall_distance.f90:75-133 on the device against the same pipeline through the oracle (laplacian + IC(0)-CG + owner values into the
    non-wall boundary slots + grad_gauss + the distance formula): bit-identical, same iteration count."""
    m = allmeshes[name]
    n = m.numCells
    ctx = make_ctx(m)
    rep = ctx.wall_distance()
    c = orc.Csr(m)
    su = np.zeros(n)
    phi = np.zeros(m.numTotal)
    a = orc.laplacian(m, c, np.ones(m.numTotal), phi, su)
    q = -m.vol[:n].copy()
    orep = orc.solve(orc.ICCG, c.ia, c.ja, a, c.diag, phi, q, 500, 1e-12, 1e-10, orc.SUM_TREE)
    for ib in range(m.numBoundaries):
        if m.bctype[ib] != M.BC_WALL:
            pf = m.patch_faces(ib)
            phi[n + pf - m.numInnerFaces] = phi[m.owner[pf] - 1]
    g = orc.grad_gauss(m, phi)[:n]
    gg = g[:, 0] * g[:, 0] + g[:, 1] * g[:, 1] + g[:, 2] * g[:, 2]
    wd = -np.sqrt(gg) + np.sqrt(gg + 2 * phi[:n])
    assert (rep.iters, rep.res0, rep.resl) == (orep.iters, orep.res0, orep.resl)
    eq(ctx.download("WALLDIST")[:n], wd, "wall distance")
    assert wd.min() > 0
    ctx.close()

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
    if name.startswith("poly"):
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


def test_piso_rejects_periodic(fcp):
    xs = np.linspace(0, 1, 5)
    m = M.hex_mesh(xs, xs, xs, dict(left="periodic", right="periodic"))
    ctx = make_ctx(m)
    with pytest.raises(L.FcpError):
        ctx.calcp_piso()
    ctx.close()

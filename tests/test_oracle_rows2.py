"""CPU: properties of the oracle's restatement of rows a9 (calcp_piso), a12 (QR gradient) and a13 (slope limiters).
The reference's tests carry no known answers for these; what can be pinned are analytic identities."""
import numpy as np
import pytest

import cases
from fcb200 import mesh as M


def test_qr_gradient_exact_for_linear_fields(orc):
    """Thin-QR least squares reproduces any linear field exactly, on orthogonal and distorted hexahedra, and agrees with
    the (correct-mode) normal-equation least squares of grad_lsq."""
    for m in (cases.golden_mesh(), M.cavity_mesh(7, distort=0.25)):
        psi = m.boundary_values_of(lambda x, y, z: 0.5 + 2 * x - 3 * y + 0.25 * z)
        D = orc.create_matrix_lsq_qr(m)
        g = orc.grad_lsq_qr(m, D, psi)
        np.testing.assert_allclose(g[: m.numCells], np.tile([2.0, -3.0, 0.25], (m.numCells, 1)), rtol=0, atol=1e-10)
        f = cases.fields(m)["p"]
        g_qr = orc.grad_lsq_qr(m, D, f)
        g_ne = orc.grad_lsq(m, False, orc.create_matrix_lsq(m, False), f, row2_correct=True)
        # (the normal-equation variant adds `small` = 1e-20 to a determinant that is ~1e-15 on the reference's 0.1 x 0.1 x 0.01
        #  mesh, gradients.f90:748: 2e-7 relative there)
        np.testing.assert_allclose(g_qr[: m.numCells], g_ne[: m.numCells], rtol=1e-6, atol=1e-6)


@pytest.mark.parametrize("kind", [1, 2, 3, 4])
def test_limiters_bounded_and_idle_on_smooth_extrema_free_data(orc, kind):
    m = M.cavity_mesh(8, distort=0.1)
    c = orc.Csr(m)
    n = m.numCells
    phi = m.boundary_values_of(lambda x, y, z: np.tanh(6 * (x - 0.5)) + 0.2 * y)
    g = orc.grad_gauss(m, phi)
    gl = orc.slope_limiter(m, c, kind, phi, g.copy())
    if kind != 4:
        # a scalar factor in [0,1] per cell
        num = (gl[:n] * g[:n]).sum(1); den = (g[:n] * g[:n]).sum(1)
        fac = num / np.where(den > 0, den, 1)
        assert np.all(fac <= 1 + 1e-12) and np.all(fac >= -1e-12)
        np.testing.assert_allclose(gl[:n], fac[:, None] * g[:n], atol=1e-12)
    else:
        # MDL: reconstructed face values never leave the neighbourhood's range
        assert np.all(np.linalg.norm(gl[:n], axis=1) <= np.linalg.norm(g[:n], axis=1) + 1e-12)
    # a linear field inside its own global range: the boundary cells hold the extrema, interior gradients survive BJ
    lin = m.boundary_values_of(lambda x, y, z: x)
    g = orc.grad_gauss(m, lin)
    gl = orc.slope_limiter(m, c, 1, lin, g.copy())
    interior = (m.xc[:n] > 0.3) & (m.xc[:n] < 0.7)
    np.testing.assert_allclose(gl[:n][interior], g[:n][interior], atol=1e-12)


def test_piso_continuity_and_reference_identities(orc):
    """After one PISO corrector solved tightly on a closed cavity: (1) sum(su) = 0 for the assembled pressure equation
    (closed domain), (2) the corrected face fluxes are discretely divergence free, (3) pp has the solver's answer and
    p = pp - mean(pp) when urfP = 1 (calcp_piso.f90:330-333)."""
    m = M.cavity_mesh(8, distort=0.15)
    f = cases.fields(m)
    c = orc.Csr(m)
    n = m.numCells
    rng = np.random.default_rng(2)
    a = -np.abs(rng.standard_normal(c.nnz)) - 0.1
    row = np.repeat(np.arange(n), np.diff(c.ia))
    a[c.diag - 1] = 1.5 * (np.bincount(row, weights=np.abs(a), minlength=n) - np.abs(a[c.diag - 1])) + 0.5
    ap = np.ones(m.numTotal); ap[:n] = 1.0 / a[c.diag - 1]
    r = [rng.standard_normal(m.numTotal) for _ in range(3)]
    g = {k: v.copy() for k, v in f.items()}
    dP = np.zeros((m.numTotal, 3)); flm = np.zeros(m.numFaces)
    reps, su, sv, sw, h = orc.calcp_piso(m, c, orc.ICCG, 2000, 1e-30, 1e-13, orc.SUM_SEQ, 1, 1, 0, 1.0, True, 0.0,
                                         r[0], r[1], r[2], g["den"], ap, ap, ap, a, g["u"], g["v"], g["w"], g["p"], g["pp"], dP, flm)
    assert reps[0].iters > 3
    np.testing.assert_array_equal(h[c.diag - 1] > 0, True)
    # (2) divergence of the corrected fluxes
    Fi = m.numInnerFaces
    div = np.zeros(n)
    np.add.at(div, m.owner[:Fi] - 1, flm[:Fi]); np.add.at(div, m.neighbour - 1, -flm[:Fi])
    assert np.abs(div).max() < 1e-9 * max(np.abs(flm).max(), 1e-30)
    # (3)
    np.testing.assert_allclose(g["p"][:n], g["pp"][:n] - g["pp"][:n].mean(), atol=1e-12 * np.abs(g["pp"]).max())


def test_face_value_family_identities(orc):
    """interpolation.f90:28-650: on a uniform orthogonal mesh with a LINEAR field and its exact gradient every second-order
    scheme returns the exact face value, every TVD limiter sees r = 1 (psi = 1 -> central), and cds returns the lambda blend."""
    m = M.cavity_mesh(6)
    n = m.numCells
    phi = m.boundary_values_of(lambda x, y, z: 1.0 + 2 * x - y + 0.5 * z)
    g = np.zeros((m.numTotal, 3)); g[:, 0], g[:, 1], g[:, 2] = 2.0, -1.0, 0.5
    f = int(m.numInnerFaces // 2)
    ijp, ijn = int(m.owner[f]), int(m.neighbour[f])
    exact = 1.0 + 2 * m.xf[f] - m.yf[f] + 0.5 * m.zf[f]
    lam = float(m.facint[f])
    for cs, name in enumerate(orc.CSCHEMES):
        vf = orc.face_value(m, cs, ijp, ijn, m.xf[f], m.yf[f], m.zf[f], 1.0 - lam, phi, g)
        if name in ("smart", "avl-smart", "boundedCentral"):   # psi(1) = 1 for all of them as well
            pass
        tol = 1e-7 if name in ("cui", "spl13", "kappa") else 1e-12      # single-precision 2./3., 1./3. literals (quirk Q5)
        assert abs(vf - exact) < tol, (name, vf, exact)


def test_calcuvw_steady_diffusion_limit(orc):
    """Pure diffusion (zero mass fluxes, uniform viscosity, no pressure gradient) on a closed cavity with a moving lid: the
    assembled momentum matrix is symmetric with negative off-diagonals and a dominant diagonal (velocity.f90:793-800,
    602-620), and the solve moves u toward the lid velocity near the top wall."""
    m = M.cavity_mesh(8)
    c = orc.Csr(m)
    n, nT = m.numCells, m.numTotal
    g = dict(u=np.zeros(nT), v=np.zeros(nT), w=np.zeros(nT), p=np.zeros(nT), den=np.ones(nT), vis=np.full(nT, 0.01), apu=np.zeros(nT),
             visw=np.full(m.numBoundaryFaces, 0.01), flmass=np.zeros(m.numFaces))
    lid = n + m.patch_faces(0) - m.numInnerFaces          # patch 0 = 'top'
    g["u"][lid] = 1.0
    prm = orc.OrcUvwParams()
    prm.solver, prm.maxiter, prm.tol_abs, prm.tol_rel = orc.BICGSTAB, 200, 1e-30, 1e-10
    prm.urf[0] = prm.urf[1] = prm.urf[2] = 1.0
    prm.gds, prm.cscheme, prm.viscos = 1.0, 0, 0.01
    a = np.zeros(c.nnz)
    o = orc.calcuvw(m, c, prm, g, a)
    import scipy.sparse as sp
    A = sp.csr_matrix((a, c.ja - 1, c.ia - 1), shape=(n, n))
    assert abs(A - A.T).max() < 1e-12
    off = A - sp.diags(A.diagonal())
    assert off.max() <= 0 and np.all(A.diagonal() >= -np.asarray(off.sum(1)).ravel() - 1e-12)
    top = m.yc[:n] > 1.0 - 1.0 / 8
    assert g["u"][:n][top].mean() > 0.3      # (v, w pick up the explicit transposed-gradient term of :805-807: not zero discretely)
    assert o["reps"][0].iters > 1

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

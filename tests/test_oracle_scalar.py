"""CPU: the oracle's row-f4 restatement (scalar_fluxes.f90 in the calcsc template of k_epsilon_rlzb.f90, modify_mu_eff) against
answers that do not depend on it: a linear diffusion profile, preservation of a constant by every convection scheme on a divergence-free
flux field, the closed form of the realizable C_mu in pure shear, the log-law wall viscosity.  The reference carries no test for these."""
import numpy as np
import pytest

import cases
import fcb200  # noqa: F401
from fcb200 import lib as L
from fcb200 import mesh as M


def params(orc, **kw):
    prm = orc.OrcScalarParams()
    prm.kind, prm.solver, prm.maxiter, prm.cscheme, prm.grad_method, prm.limiter, prm.tscheme, prm.sum_mode = 0, 3, 500, 0, 0, 0, 0, 0
    prm.tol_abs, prm.tol_rel, prm.urf, prm.gds, prm.timestep, prm.prtr, prm.viscos, prm.densit = 1e-30, 1e-13, 1.0, 1.0, 0.0, 1.0, 0.01, 1.0
    for k, v in kw.items():
        setattr(prm, k, v)
    return prm


def channel(distort=0.0):
    return M.hex_mesh(np.linspace(0, 2.0, 17), np.linspace(0, 1.0, 7), np.linspace(0, 0.5, 5),
                      dict(left="inlet", right="outlet", back="symmetry", front="symmetry"), distort=distort)


def test_pure_diffusion_gives_the_linear_profile(orc):
    """No flow, phi = 1 on the inlet patch, 0 on the outlet patch, zero-flux walls: phi = 1 - x/L exactly on an orthogonal mesh."""
    m = channel()
    c = orc.Csr(m)
    n = m.numCells
    f = dict(phi=np.zeros(m.numTotal), den=np.ones(m.numTotal), vis=np.full(m.numTotal, 0.03), flmass=np.zeros(m.numFaces),
             su_vol=np.zeros(n), sp_vol=np.zeros(n))
    for ib in range(m.numBoundaries):
        if m.bctype[ib] == M.BC_INLET:
            f["phi"][n + m.patch_faces(ib) - m.numInnerFaces] = 1.0
    inlet_outlet = np.concatenate([n + m.patch_faces(ib) - m.numInnerFaces for ib in range(m.numBoundaries) if m.bctype[ib] in (M.BC_INLET, M.BC_OUTLET)])
    keep = f["phi"][inlet_outlet].copy()
    for _ in range(3):                       # updateBoundary copies the owner value into the outlet slots: re-impose the Dirichlet data
        orc.calcsc(m, c, params(orc, tol_abs=1e-13, tol_rel=1e-10), f)
        f["phi"][inlet_outlet] = keep
    assert np.abs(f["phi"][:n] - (1.0 - m.xc[:n] / 2.0)).max() < 1e-10


@pytest.mark.parametrize("cscheme", ["cds", "central", "linearUpwind", "muscl", "vanleer", "quick"])
def test_a_constant_is_preserved_by_every_scheme(orc, cscheme):
    """Uniform flow (divergence-free face fluxes), phi = 3 everywhere including the inlet: the discrete equation must return phi = 3
    (signs of can/cap, the upwind split ce/cp, the deferred correction and the boundary coefficients all have to be consistent)."""
    m = channel(distort=0.15)
    c = orc.Csr(m)
    n, Fi = m.numCells, m.numInnerFaces
    flm = 1.2 * m.arx.copy()                 # rho u . S with u = (1,0,0), rho = 1.2
    for ib in range(m.numBoundaries):
        if m.bctype[ib] in (M.BC_WALL, M.BC_SYMMETRY):
            flm[m.patch_faces(ib)] = 0.0
    div = np.zeros(n)
    np.add.at(div, m.owner[:Fi] - 1, flm[:Fi]); np.add.at(div, m.neighbour - 1, -flm[:Fi]); np.add.at(div, m.owner[Fi:] - 1, flm[Fi:])
    assert np.abs(div).max() < 1e-14
    f = dict(phi=np.full(m.numTotal, 3.0), den=np.full(m.numTotal, 1.2), vis=np.full(m.numTotal, 0.02), flmass=flm, su_vol=np.zeros(n), sp_vol=np.zeros(n))
    o = orc.calcsc(m, c, params(orc, cscheme=L.CSCHEME_ID[cscheme], gds=0.9, urf=0.8), f)
    assert np.abs(f["phi"][:n] - 3.0).max() < 1e-10, cscheme
    assert (o["fimin"], o["fimax"]) == (f["phi"][:n].min(), f["phi"][:n].max())


def test_realizable_cmu_in_pure_shear(orc):
    """u = S y: S_ij S_jk S_ki = 0 -> phi = acos(0)/3 = pi/6, A_s = sqrt(6) cos(pi/6), U* = S, C_mu = 1/(A0 + A_s S k/eps) (Shih et al. 1995)."""
    m = M.cavity_mesh(6)
    n = m.numCells
    S, k, eps, rho, nu = 3.0, 0.04, 0.5, 1.1, 1e-3
    gU = np.zeros((m.numTotal, 3)); gU[:, 1] = S
    z3 = np.zeros((m.numTotal, 3))
    te, ed, den = np.full(m.numTotal, k), np.full(m.numTotal, eps), np.full(m.numTotal, rho)
    u = m.boundary_values_of(lambda x, y, z: S * y); v = np.zeros(m.numTotal); w = np.zeros(m.numTotal)
    vis = np.full(m.numTotal, nu); visw = np.zeros(m.numBoundaryFaces)
    dnw = np.full(m.numBoundaryFaces, 0.05)
    orc.modify_mu_eff_rlzb(m, 1.0, nu, gU, z3, z3, te, ed, den, u, v, w, dnw, vis, visw)
    # the reference writes `1./3.` and `a0 = 4.04` as default-real literals (quirk Q5): phi = float32(1/3) * acos(0), not exactly pi/6
    ffi = float(np.float32(1.0) / np.float32(3.0)) * np.arccos(0.0)
    assert abs(ffi - np.pi / 6) < 1e-7
    cmu = 1.0 / (float(np.float32(4.04)) + np.sqrt(6.0) * np.cos(ffi) * S * k / eps)
    assert np.abs(vis[:n] - (nu + rho * cmu * k * k / eps)).max() < 1e-14
    # strain invariants of the same field
    ms, vort = orc.calc_strain_and_vorticity(m, gU, z3, z3)
    assert np.allclose(ms, S, rtol=0, atol=1e-15) and np.allclose(vort, S, rtol=0, atol=1e-15)


def test_wall_function_viscosity(orc):
    """y* = rho cmu^1/4 sqrt(k) d / mu; above 11.63 the wall viscosity is mu y* kappa / ln(E y*), below it the molecular one."""
    m = M.cavity_mesh(4)
    n = m.numCells
    z3 = np.zeros((m.numTotal, 3))
    nu, rho = 1e-3, 1.0
    for k, expect_log in ((1e-8, False), (0.5, True)):
        te, ed, den = np.full(m.numTotal, k), np.full(m.numTotal, 1.0), np.full(m.numTotal, rho)
        u = m.boundary_values_of(lambda x, y, z: 1.0 + 0 * x); v = np.zeros(m.numTotal); w = np.zeros(m.numTotal)
        vis = np.full(m.numTotal, nu); visw = np.zeros(m.numBoundaryFaces); dnw = np.full(m.numBoundaryFaces, 0.125)
        ypl, tau = orc.modify_mu_eff_rlzb(m, 1.0, nu, z3, z3, z3, te, ed, den, u, v, w, dnw, vis, visw)
        ystar = rho * 0.09 ** 0.25 * np.sqrt(k) * 0.125 / nu
        assert np.allclose(ypl[: m.numBoundaryFaces], ystar, rtol=1e-14)
        want = nu * ystar * 0.41 / np.log(8.432 * ystar) if expect_log else nu
        assert (ystar > 11.63) == expect_log
        assert np.allclose(visw, want, rtol=1e-13) and np.allclose(vis[n:], want, rtol=1e-13)


def test_epsilon_wall_cells_get_the_equilibrium_value(orc):
    m = cases.meshes()["hex10_distorted"]
    import test_gpu_scalar as T
    g = T.scalar_inputs(m, orc)
    c = orc.Csr(m)
    prm = params(orc, kind=orc.SC_EPS_RLZB, maxiter=50, tol_rel=1e-10, urf=1.0, prtr=1.0 / 1.2)
    te0 = g["te"].copy()
    o = orc.calcsc(m, c, prm, g)
    n, Fi = m.numCells, m.numInnerFaces
    # a cell with one wall face holds cmu^0.75 k^1.5 / (kappa dnw) after the solve (identity row, sp = 1, su = ed)
    nwall = np.zeros(n, int)
    np.add.at(nwall, m.owner[Fi:] - 1, 1)
    one = np.nonzero(nwall == 1)[0]
    bf_of = {int(m.owner[Fi + b] - 1): b for b in range(m.numBoundaryFaces)}
    for cell in one[:20]:
        want = 0.09 ** 0.75 * te0[cell] ** 1.5 / (0.41 * g["dnw"][bf_of[int(cell)]])
        assert abs(g["ed"][cell] - want) <= 1e-9 * want


def test_fvx_gauss_gradient_is_exact_for_linear_fields_on_skewed_meshes(orc):
    """The second pass of fvxGradient's grad_gauss corrects the face value for skewness: on a distorted mesh the plain Gauss gradient of a
    linear field is only approximate, the two-pass one is much closer (and exact on an orthogonal mesh)."""
    lin = lambda x, y, z: 2 * x - 3 * y + 0.5 * z  # noqa: E731
    m = M.cavity_mesh(8)
    gx, gy, gz = orc.grad_gauss_fvx(m, m.boundary_values_of(lin))
    assert np.abs(gx - 2).max() < 1e-12 and np.abs(gy + 3).max() < 1e-12 and np.abs(gz - 0.5).max() < 1e-12
    m = M.cavity_mesh(8, distort=0.25)
    phi = m.boundary_values_of(lin)
    gx, gy, gz = orc.grad_gauss_fvx(m, phi)
    plain = orc.grad_gauss(m, phi)[: m.numCells]
    err2 = max(np.abs(gx - 2).max(), np.abs(gy + 3).max(), np.abs(gz - 0.5).max())
    err1 = np.abs(plain - np.array([2.0, -3.0, 0.5])).max()
    assert err2 < 0.5 * err1, (err1, err2)


def test_sgs_models_in_pure_shear(orc):
    """u = S y on a uniform mesh.  WALE: the traceless symmetric part of the squared gradient vanishes in pure shear, so mu_sgs = 0 (the
    design property of the model, Nicoud & Ducros 1999).  Vreman: alpha_ij = d_j u_i has one entry, beta = alpha^T alpha has one diagonal entry,
    B_beta = 0, so mu_sgs = 0 too (Vreman 2004, section II) -- both return the laminar viscosity; in the interior the reference's inner-product
    slip (quirk Q24) does not touch these components."""
    m = M.cavity_mesh(6)
    n = m.numCells
    S, nu = 3.0, 1e-3
    u = m.boundary_values_of(lambda x, y, z: S * y); v = np.zeros(m.numTotal); w = np.zeros(m.numTotal)
    for model in (orc.SGS_WALE, orc.SGS_VREMAN):
        vis = np.full(m.numTotal, 0.5); visw = np.zeros(m.numBoundaryFaces)
        orc.modify_viscosity_sgs(m, model, 1.0, nu, u, v, w, np.ones(m.numTotal), vis, visw)
        assert np.abs(vis[:n] - nu).max() < 1e-12, model
        assert np.all(visw == nu)


def test_sst_closed_forms(orc):
    """k-omega SST (Menter 1994 / 2003) on uniform fields: with S = 0 the eddy viscosity is rho k / omega; with zero gradients the third argument
    of F1 drops out and F1 = tanh(max(sqrt(k)/(beta* d omega), 500 nu/(d^2 omega))^4), near 1 at the wall and near 0 far from it; the omega
    imposed in wall cells is sqrt((6 nu/(beta1 d^2))^2 + (sqrt(k)/(cmu^1/4 kappa d))^2)."""
    m = M.cavity_mesh(6)
    n, nT, B = m.numCells, m.numTotal, m.numBoundaryFaces
    c = orc.Csr(m)
    k, om, rho, nu = 0.02, 30.0, 1.0, 1e-4
    d = np.zeros(nT); d[:n] = np.minimum.reduce([m.xc[:n], 1 - m.xc[:n], m.yc[:n], 1 - m.yc[:n], m.zc[:n], 1 - m.zc[:n]])
    te, ed, den = np.full(nT, k), np.full(nT, om), np.full(nT, rho)
    vis = np.full(nT, nu); visw = np.zeros(B)
    u = np.full(nT, 0.3); v = np.zeros(nT); w = np.zeros(nT)
    orc.modify_mu_eff_sst(m, 1.0, nu, rho, False, np.zeros(n), d, te, ed, den, u, v, w, np.full(B, 0.08), vis, visw)
    assert np.abs(vis[:n] - (nu + rho * k / om)).max() < 1e-15
    prm = params(orc, kind=orc.SC_OMEGA_SST, maxiter=1, viscos=nu, densit=rho)
    f = dict(te=te, ed=ed.copy(), den=den, vis=np.full(nT, nu + rho * k / om), visw=np.full(B, nu), dnw=np.full(B, 0.08), flmass=np.zeros(m.numFaces),
             u=u, v=v, w=w, magStrain=np.zeros(n), gen=np.zeros(n), fsst=np.zeros(nT), walldist=d, dTEdxi=np.zeros((nT, 3)))
    orc.calcsc(m, c, prm, f)
    ksi = np.maximum(np.sqrt(k) / (0.09 * d[:n] * om + 1e-20), 500 * nu / rho / (d[:n] ** 2 * om + 1e-20))
    assert np.allclose(f["fsst"][:n], np.tanh(ksi ** 4), rtol=1e-13, atol=0)
    near, far = d[:n] < 0.1, d[:n] > 0.4
    assert f["fsst"][:n][near].min() > 0.1 > f["fsst"][:n][far].max()
    # wall cells hold the imposed omega after the (identity-row) solve
    Fi = m.numInnerFaces
    nwall = np.zeros(n, int); np.add.at(nwall, m.owner[Fi:] - 1, 1)
    want = np.sqrt((6 * nu / rho / (0.075 * 0.08 ** 2)) ** 2 + (np.sqrt(k) / (0.09 ** 0.25 * 0.41 * 0.08)) ** 2)
    assert np.allclose(f["ed"][:n][nwall >= 1], want, rtol=1e-12)


def test_iterative_gauss_gradient_of_the_mpi_tree(orc):
    """src-par/gradients.f90:1547-1664 (nigrad passes of gradco): exact for a linear field on an orthogonal mesh for any pass count; on a skewed mesh
    every further pass brings the gradient of a linear field closer to the exact one (the passes are a fixed-point iteration of the skewness
    correction); nigrad = 2 is the fvx gradient bit for bit."""
    import numpy as np
    from fcb200 import mesh as M
    lin = lambda x, y, z: 0.5 + 2.0 * x - 1.5 * y + 0.75 * z      # noqa: E731
    exact = np.array([2.0, -1.5, 0.75])
    m = M.cavity_mesh(8, bump=0.4)
    for k in (1, 2, 3):
        g = np.stack(orc.grad_gauss_iter(m, m.boundary_values_of(lin), k), 1)
        assert np.abs(g - exact).max() < 1e-12
    ms = M.cavity_mesh(8, distort=0.25)
    phi = ms.boundary_values_of(lin)
    err = [np.abs(np.stack(orc.grad_gauss_iter(ms, phi, k), 1) - exact).max() for k in (1, 2, 3, 4)]
    assert err[1] < 0.5 * err[0] and err[2] < err[1] and err[3] < err[2], err
    a = orc.grad_gauss_iter(ms, phi, 2)
    b = orc.grad_gauss_fvx(ms, phi)
    assert all(np.array_equal(x, y) for x, y in zip(a, b))

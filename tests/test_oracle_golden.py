"""CPU: pins the oracle (oracle/orc.cpp) against the known answers the reference's own tests carry
 and against the committed regression pins."""
import json
import math
import os

import numpy as np
import pytest

import cases
from fcb200 import mesh as M


@pytest.fixture(scope="module")
def gold():
    with open(os.path.join(cases.GOLDEN, "spsolve_5x5.json")) as fh:
        return json.load(fh)


@pytest.mark.parametrize("system,solver", [("nonsymmetric", "bicgstab"), ("spd", "iccg"), ("spd", "dpcg")])
@pytest.mark.parametrize("f32", [False, True])
def test_5x5_systems(orc, gold, system, solver, f32):
    """test/test_linear_solvers_spsolve.f90: printed analytic solutions, two decimals."""
    g = gold[system]
    a = np.array(g["a_f32" if f32 else "a"]); b = np.array(g["b_f32" if f32 else "b"])
    ia = np.array(gold["ioffset"], np.int32); ja = np.array(gold["ja"], np.int32); diag = np.array(gold["diag"], np.int32)
    x = np.zeros(5)
    sid = {"dpcg": orc.DPCG, "iccg": orc.ICCG, "bicgstab": orc.BICGSTAB}[solver]
    # the reference test uses itr_max = 5 (exact convergence of a Krylov method on a 5x5 system in <= 5 steps);
    # dpcg/iccg on the (ill-conditioned) SPD system need a few more in floating point
    rep = orc.solve(sid, ia, ja, a, diag, x, b, 50, 1e-10, 1e-7)
    assert rep.iters <= 50
    np.testing.assert_allclose(x, np.array(g["x"]), atol=0.0051 + 2e-3 * np.abs(np.array(g["x"])).max())
    # residual really small
    r = b - a.reshape(5, 5) @ x
    assert np.abs(r).sum() < 1e-6 * np.abs(b).sum()


def test_gauss_seidel_on_the_spd_5x5_system(orc, gold):
    """The same SPD system through the oracle's Gauss-Seidel (linear_solvers.f90:96-201): it converges to the printed solution, the residual
    norm falls monotonically (SPD matrix) and the early return is taken AFTER the first sweep has updated fi (:139-157)."""
    g = gold["spd"]
    a = np.array(g["a"]); b = np.array(g["b"])
    ia = np.array(gold["ioffset"], np.int32); ja = np.array(gold["ja"], np.int32); diag = np.array(gold["diag"], np.int32)
    x = np.zeros(5)
    rep = orc.solve(orc.GAUSS_SEIDEL, ia, ja, a, diag, x, b, 5000, 1e-13, 1e-11)
    assert 1 < rep.iters < 5000
    assert np.allclose(x, g["x"], atol=0.0051)
    norms = []
    for its in range(1, 6):
        y = np.zeros(5)
        norms.append(orc.solve(orc.GAUSS_SEIDEL, ia, ja, a, diag, y, b, its, 1e-300, 1e-300).resl)
    assert all(n1 < n0 for n0, n1 in zip(norms, norms[1:]))
    y = np.zeros(5)
    r1 = orc.solve(orc.GAUSS_SEIDEL, ia, ja, a, diag, y, b, 50, 1e30, 1e-11)
    assert r1.iters == 1 and np.abs(y).max() > 0 and "No Iterations 1" in orc.report_line(orc.GAUSS_SEIDEL, "x", r1)
    # the first unknown of the first sweep by hand: fi(1) = 0 + (b(1) - 0)/(a(diag(1)) + small)
    assert y[0] == b[0] / (a[diag[0] - 1] + np.float64(np.float32(1e-20)))


def test_gauss_gradient_of_linear_field_is_111(orc):
    """test/testFieldOperations/testFieldOperations.f90:137-160 on the reference's own mesh."""
    m = cases.golden_mesh()
    psi = m.boundary_values_of(lambda x, y, z: x + y + z)
    g = orc.grad_gauss(m, psi)
    np.testing.assert_allclose(g[: m.numCells], 1.0, rtol=0, atol=2e-12)
    assert np.all(g[m.numCells:] == 0.0)


def test_lsq_gradient_of_linear_field(orc):
    """Our own identity (SURVEY 8c item 3 caveat): rows 1 and 3 exact; row 2 exact only in 'correct' mode (quirk Q1)."""
    m = M.cavity_mesh(8, distort=0.2)
    psi = m.boundary_values_of(lambda x, y, z: 2 * x - 3 * y + 0.5 * z)
    for w in (False, True):
        D = orc.create_matrix_lsq(m, w)
        g = orc.grad_lsq(m, w, D, psi, row2_correct=True)
        if not w:   # weighted variant carries quirk Q2 at boundary cells
            np.testing.assert_allclose(g[: m.numCells], np.broadcast_to(np.array([2.0, -3.0, 0.5]), (m.numCells, 3)), atol=1e-10)
        gr = orc.grad_lsq(m, w, D, psi, row2_correct=False)
        np.testing.assert_array_equal(gr[:, 0], g[:, 0])
        np.testing.assert_array_equal(gr[:, 2], g[:, 2])


def test_q1_value_on_reference_mesh(orc):
    """SURVEY: with bug Q1 the reference's LSQ y-component equals d11/d22 on orthogonal cells; on the 20x20x1 test mesh
    (dx = dy = 0.05 m/20..., dz = one layer) interior cells still give 1."""
    m = cases.golden_mesh()
    psi = m.boundary_values_of(lambda x, y, z: x + y + z)
    D = orc.create_matrix_lsq(m, False)
    g = orc.grad_lsq(m, False, D, psi, row2_correct=False)
    interior = np.ones(m.numCells, bool)
    interior[m.owner[m.numInnerFaces:] - 1] = False
    np.testing.assert_allclose(g[: m.numCells][interior], 1.0, atol=1e-10)


def test_poisson_known_answer_second_order(orc):
    """applications/Poisson/poisson.f90:63-104: -lap(p) = 8 pi^2 sin(2 pi x) sin(2 pi y); laplacian(-1, p) makes every
    patch Dirichlet, so the exact (z-independent) solution is imposed on all six patches; L_inf error is O(h^2)."""
    errs = []
    pi = np.pi
    exact = lambda x, y, z: np.sin(2 * pi * x) * np.sin(2 * pi * y)  # noqa: E731
    for n in (8, 16, 32):
        m = M.hex_mesh(np.linspace(0, 1, n + 1), np.linspace(0, 1, n + 1), np.linspace(0, 4.0 / n, 5))
        csr = orc.Csr(m)
        nC = m.numCells
        phi = m.boundary_values_of(exact)
        phi[:nC] = 0.0
        su = 8 * pi * pi * exact(m.xc[:nC], m.yc[:nC], 0) * m.vol[:nC]
        a = orc.laplacian(m, csr, -np.ones(m.numTotal), phi, su)
        x = np.zeros(nC)
        rep = orc.solve(orc.ICCG, csr.ia, csr.ja, a, csr.diag, x, su, 1000, 1e-30, 1e-13)
        assert 0 < rep.iters < 1000
        errs.append(np.abs(x - exact(m.xc[:nC], m.yc[:nC], 0)).max())
    assert errs[1] < errs[0] / 3.0 and errs[2] < errs[1] / 3.5, errs


@pytest.mark.parametrize("solver", ["dpcg", "iccg"])
def test_poisson_app_reference_setup(orc, solver):
    """The reference's exact setup (all boundary values 0): both PCG variants reach the same discrete solution."""
    m = M.hex_mesh(np.linspace(0, 1, 21), np.linspace(0, 1, 21), np.array([0.0, 0.05]))
    csr, a, su = cases.poisson_system(m, orc)
    x = np.zeros(m.numCells)
    sid = {"dpcg": orc.DPCG, "iccg": orc.ICCG}[solver]
    rep = orc.solve(sid, csr.ia, csr.ja, a, csr.diag, x, su, 1000, 1e-30, 1e-12)
    assert 0 < rep.iters < 1000
    r = su - orc.spmv(csr.ia, csr.ja, a, x)
    assert np.abs(r).sum() <= 1e-11 * np.abs(su).sum()


def test_wall_distance_pipeline(orc):
    """src/mesh/wall_distance.f90:96-133: laplacian(1,phi) + iccg(500, 1e-12, 1e-10) + grad_gauss -> distance to the nearest wall."""
    n = 24
    m = M.cavity_mesh(n)
    csr = orc.Csr(m)
    nC = m.numCells
    su = -m.vol[:nC].copy()
    phi = np.zeros(m.numTotal)
    a = orc.laplacian(m, csr, np.ones(m.numTotal), phi, su)
    x = np.zeros(m.numTotal)
    rep = orc.solve(orc.ICCG, csr.ia, csr.ja, a, csr.diag, x, su, 500, 1e-12, 1e-10)
    assert 0 < rep.iters < 500
    g = orc.grad_gauss(m, x)[:nC]
    gm = np.sqrt((g * g).sum(1))
    d = -gm + np.sqrt(gm * gm + 2 * x[:nC])
    exact = np.minimum.reduce([m.xc[:nC], 1 - m.xc[:nC], m.yc[:nC], 1 - m.yc[:nC], m.zc[:nC], 1 - m.zc[:nC]])
    near = exact < 0.1
    assert np.abs(d[near] - exact[near]).max() < 0.03


@pytest.mark.parametrize("name", ["hex6", "hex10_distorted", "slab39_empty"])
def test_closed_cavity_sum_su_is_zero(orc, name):
    """Pressure/calcp_simple.f90:314: sum(su) = 0 for a closed domain; also A is symmetric with zero row sums."""
    m = cases.meshes()[name]
    f = cases.fields(m)
    csr = orc.Csr(m)
    dP = orc.grad_gauss(m, f["p"])
    a, su, flm = orc.assemble_pcorr(m, csr, f["den"], f["u"], f["v"], f["w"], f["p"], f["pp"], dP, f["apu"])
    assert abs(su.sum()) <= 1e-12 * np.abs(flm).sum()
    rows = np.repeat(np.arange(m.numCells), np.diff(csr.ia))
    rs = np.bincount(rows, weights=a, minlength=m.numCells)
    assert np.abs(rs).max() <= 1e-12 * np.abs(a).max()
    assert np.array_equal(a[csr.icell_jcell - 1], a[csr.jcell_icell - 1])


def test_oracle_regression_pins(orc):
    pins = np.load(os.path.join(cases.GOLDEN, "oracle_pins.npz"))
    m = cases.golden_mesh()
    csr = orc.Csr(m)
    for k, v in (("ia", csr.ia), ("ja", csr.ja), ("diag", csr.diag), ("kpn", csr.icell_jcell), ("knp", csr.jcell_icell)):
        np.testing.assert_array_equal(pins[k], v)
    phi = pins["phi"]
    np.testing.assert_array_equal(pins["grad_gauss"], orc.grad_gauss(m, phi))
    for w, nm in ((False, "lsq"), (True, "lsq_dm")):
        D = orc.create_matrix_lsq(m, w)
        np.testing.assert_array_equal(pins["Dmat_" + nm], D)
        np.testing.assert_array_equal(pins["grad_" + nm], orc.grad_lsq(m, w, D, phi))
    su = np.zeros(m.numCells)
    np.testing.assert_array_equal(pins["lap_a"], orc.laplacian(m, csr, np.ones(m.numTotal), phi, su))
    np.testing.assert_array_equal(pins["lap_su"], su)


def test_csr_pattern_properties(orc):
    """sparse_matrix.f90:110-260: nnz = N + 2F, rows sorted, diagonal embedded, face maps point at the right columns."""
    for name, m in cases.meshes().items():
        csr = orc.Csr(m)
        assert csr.nnz == m.numCells + 2 * (m.numInnerFaces + m.numPeriodic)        # :110
        assert csr.ia[0] == 1 and csr.ia[-1] == csr.nnz + 1
        for i in range(0, m.numCells, max(1, m.numCells // 50)):
            row = csr.ja[csr.ia[i] - 1: csr.ia[i + 1] - 1]
            assert np.all(np.diff(row) > 0)
            assert csr.ja[csr.diag[i] - 1] == i + 1
        Fi = m.numInnerFaces
        np.testing.assert_array_equal(csr.ja[csr.icell_jcell[:Fi] - 1], m.neighbour)
        np.testing.assert_array_equal(csr.ja[csr.jcell_icell[:Fi] - 1], m.owner[:Fi])


def test_sum_tree_matches_fsum(orc):
    rng = np.random.default_rng(1)
    for n in (0, 1, 31, 256, 2048, 2049, 100003):
        v = rng.standard_normal(n)
        assert abs(orc.sum_tree(v) - math.fsum(v)) <= 1e-13 * (np.abs(v).sum() + 1)


def test_solver_modes_agree(orc):
    """SEQ (gfortran order) and TREE (GPU order) summation give the same iteration counts and solutions to rounding."""
    m = M.cavity_mesh(12, bump=0.5)
    csr, a, su = cases.poisson_system(m, orc)
    for sid in (orc.DPCG, orc.ICCG, orc.BICGSTAB):
        xs, its = [], []
        for mode in (orc.SUM_SEQ, orc.SUM_TREE):
            x = np.zeros(m.numCells)
            rep = orc.solve(sid, csr.ia, csr.ja, a, csr.diag, x, su, 500, 1e-30, 1e-10, mode)
            xs.append(x); its.append(rep.iters)
        assert abs(its[0] - its[1]) <= 1
        np.testing.assert_allclose(xs[0], xs[1], rtol=1e-7, atol=1e-12)


def test_partitioned_dpcg_matches_serial(orc):
    """src-par layout with virtual ranks (orc_dpcg_par) vs the unpartitioned solver: same counts, same solution."""
    m = M.cavity_mesh(10)
    csr, a, su = cases.poisson_system(m, orc)
    x = np.zeros(m.numCells)
    rep = orc.solve(orc.DPCG, csr.ia, csr.ja, a, csr.diag, x, su, 500, 1e-30, 1e-10)
    for P in (2, 4):
        parts = M.partition(m, M.slab_partition(m, P))
        csrs = [orc.Csr(p) for p in parts]
        a_l, apr_l, fi_l, rhs_l = [], [], [], []
        for p, c in zip(parts, csrs):
            al, apr = M.localize_matrix(m, csr, a, p, c)
            a_l.append(al); apr_l.append(apr)
            fi_l.append(np.zeros(p.numTotal)); rhs_l.append(su[p.cell_global].copy())
        rp = orc.dpcg_par(parts, csrs, a_l, apr_l, fi_l, rhs_l, 500, 1e-30, 1e-10)
        assert abs(rp.iters - rep.iters) <= 1
        xg = np.zeros(m.numCells)
        for p, fi in zip(parts, fi_l):
            xg[p.cell_global] = fi[: p.numCells]
        np.testing.assert_allclose(xg, x, rtol=1e-6, atol=1e-10)

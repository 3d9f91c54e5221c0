"""Shared by the CPU and GPU outer-loop tests: the reference's sequential SIMPLE iteration (src/cappuccino/main.f90:142-160:
`call calcuvw` then `call calcp_simple`) on the lid-driven cavity of examples/cavity (Re = 100: lid velocity 1, L = 1,
densit = 1, viscos = 0.01; input.nml: central, gauss, urfU = 0.8, urfP = 0.3, bicgstab 5 its tolRel 0.01, iccg 20 its
tolRel 0.025, pRefCell = 1), driven through the oracle."""
import numpy as np

from fcb200 import lib as L
from fcb200 import mesh as M

# Ghia, Ghia & Shin, J. Comput. Phys. 48 (1982) 387-411, Table I: u along the vertical centre line, Re = 100
GHIA_Y = np.array([0.9766, 0.9688, 0.9609, 0.9531, 0.8516, 0.7344, 0.6172, 0.5000, 0.4531, 0.2813, 0.1719, 0.1016, 0.0703, 0.0625, 0.0547])
GHIA_U = np.array([0.84123, 0.78871, 0.73722, 0.68717, 0.23151, 0.00332, -0.13641, -0.20581, -0.21090, -0.15662, -0.10150, -0.06434, -0.04775,
                   -0.04192, -0.03717])

INPUT = dict(urfU=(0.8, 0.8, 0.8), gdsU=1.0, cSchemeU="central", lSolverU="bicgstab", maxiterU=5, tolAbsU=1e-15, tolRelU=0.01, pscheme="linear",
             urfP=0.3, lSolverP="iccg", maxiterP=20, tolAbsP=1e-15, tolRelP=0.025, viscos=0.01, pRefCell=1)


def cavity(n=39):
    """examples/cavity/cavity.geo: 40 x 40 nodes, bump 0.2, one layer in z, front/back 'empty' (README.md)."""
    return M.cavity_mesh(n, nz=1, bump=0.2)


def initial_state(m):
    nT = m.numTotal
    f = dict(u=np.zeros(nT), v=np.zeros(nT), w=np.zeros(nT), p=np.zeros(nT), pp=np.zeros(nT), den=np.ones(nT), vis=np.full(nT, INPUT["viscos"]),
             apu=np.zeros(nT), apv=np.zeros(nT), apw=np.zeros(nT), visw=np.full(m.numBoundaryFaces, INPUT["viscos"]), flmass=np.zeros(m.numFaces))
    lid = m.numCells + m.patch_faces(0) - m.numInnerFaces     # patch 0 = 'top'
    f["u"][lid] = 1.0
    return f


def oracle_params(orc, sum_mode):
    prm = orc.OrcUvwParams()
    prm.solver, prm.maxiter, prm.tol_abs, prm.tol_rel = orc.BICGSTAB, INPUT["maxiterU"], INPUT["tolAbsU"], INPUT["tolRelU"]
    prm.urf[0], prm.urf[1], prm.urf[2] = INPUT["urfU"]
    prm.gds, prm.cscheme, prm.viscos, prm.pscheme, prm.sum_mode = INPUT["gdsU"], L.CSCHEME_ID[INPUT["cSchemeU"]], INPUT["viscos"], 0, sum_mode
    return prm


def oracle_iteration(orc, m, c, prm, f, a, dP, sum_mode):
    """One outer iteration; f, a, dP are updated in place.  Returns (momentum reports, pressure report)."""
    o = orc.calcuvw(m, c, prm, f, a)
    f["apv"][:], f["apw"][:] = o["apv"], o["apw"]
    dP[...] = o["dPdxi"]
    su = np.zeros(m.numCells)
    orc.assemble_pcorr_into(m, c, f["den"], f["u"], f["v"], f["w"], f["p"], f["pp"], dP, f["apu"], a, su, f["flmass"], const_mflux=True)
    rep = orc.solve(orc.ICCG, c.ia, c.ja, a, c.diag, f["pp"], su, INPUT["maxiterP"], INPUT["tolAbsP"], INPUT["tolRelP"], sum_mode)
    orc.correct_simple(m, c, 0, a, f["den"], f["u"], f["v"], f["w"], f["p"], f["pp"], f["apu"], f["apv"], f["apw"], INPUT["urfP"], INPUT["pRefCell"],
                       dP, f["flmass"])
    return o["reps"], rep


def centreline_u(m, u):
    """u(y) on the vertical centre line x = 0.5 (the odd 39-cell mesh has a cell column exactly there)."""
    n = m.numCells
    x = m.xc[:n]
    col = np.abs(x - 0.5) < 1e-9
    assert col.sum() > 0
    order = np.argsort(m.yc[:n][col])
    return m.yc[:n][col][order], u[:n][col][order]

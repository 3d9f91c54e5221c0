"""Shared by the CPU and GPU tests of BASELINE config 2: the reference's own examples/pitzDaily case (backward-facing step of Pitz & Daily,
12 225 hexahedra, OpenFOAM polyMesh: in = inlet, out = outlet, upperWall / lowerWall = wall, sides = symmetry) with the settings of
examples/pitzDaily/input-simple.nml (muscl, Venkatakrishnan, gauss, weighted pressure interpolation, urfU 0.5, urfP 0.3, BiCGStab-ILU(0)
10 its for U, IC(0)-CG 100 its for p', realizable k-epsilon with the TurbModelData.f90 defaults: linearUpwind, urf 0.7, BiCGStab 10 its) and
the fields of its 0/ directory (U_in = 10, k = 0.375, epsilon = 14.855), driven in the order of src/cappuccino/main.f90:142-175."""
import os

import numpy as np

from fcb200 import lib as L
from fcb200 import mesh as M

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
I = dict(urfU=(0.5, 0.5, 0.5), gdsU=1.0, cSchemeU="muscl", maxiterU=10, tolAbsU=1e-13, tolRelU=0.01, pscheme="weighted", urfP=0.3, maxiterP=100,
         tolAbsP=1e-13, tolRelP=0.01, densit=1.0, viscos=1e-5, limiter="Venkatakrishnan", pRefCell=1,
         sc=dict(urf=0.7, gds=1.0, cscheme="linearUpwind", maxiter=10, tol_abs=1e-10, tol_rel=0.01), urfVis=1.0, Uin=10.0, k0=0.375, eps0=14.855)
CMU = 0.09


def mesh():
    return M.load_mesh_npz(os.path.join(GOLDEN, "pitzDaily_polymesh.npz"))


def initial_state(m, orc):
    n, nT, Fi, B = m.numCells, m.numTotal, m.numInnerFaces, m.numBoundaryFaces
    f = dict(u=np.zeros(nT), v=np.zeros(nT), w=np.zeros(nT), p=np.zeros(nT), pp=np.zeros(nT), den=np.full(nT, I["densit"]), vis=np.full(nT, I["viscos"]),
             apu=np.zeros(nT), apv=np.zeros(nT), apw=np.zeros(nT), visw=np.full(B, I["viscos"]), flmass=np.zeros(m.numFaces),
             te=np.full(nT, I["k0"]), ed=np.full(nT, I["eps0"]))
    dnw, srdw, dns, srds = orc.wall_geometry(m)
    f["dnw"] = np.full(B, 1.0)
    iw = 0
    for ib in range(m.numBoundaries):
        pf = m.patch_faces(ib); sl = n + pf - Fi
        if m.bctype[ib] == M.BC_WALL:
            f["dnw"][pf - Fi] = dnw[iw: iw + pf.size]; iw += pf.size
        if m.bctype[ib] == M.BC_INLET:
            f["u"][sl] = I["Uin"]
            f["flmass"][pf] = f["den"][sl] * (f["u"][sl] * m.arx[pf] + f["v"][sl] * m.ary[pf] + f["w"][sl] * m.arz[pf])
            f["vis"][sl] = I["viscos"] + f["den"][sl] * f["te"][sl] ** 2 * CMU / (f["ed"][sl] + 1e-20)      # modify_viscosity_inlet_k_epsilon_rlzb
    flomas = float(-f["flmass"].sum())
    return f, flomas


def oracle_params(orc, sum_mode):
    up = orc.OrcUvwParams()
    up.solver, up.maxiter, up.tol_abs, up.tol_rel = orc.BICGSTAB, I["maxiterU"], I["tolAbsU"], I["tolRelU"]
    up.urf[0], up.urf[1], up.urf[2] = I["urfU"]
    up.gds, up.cscheme, up.limiter, up.pscheme, up.viscos, up.sum_mode = I["gdsU"], L.CSCHEME_ID[I["cSchemeU"]], L.LIMITER_ID[I["limiter"]], 2, I["viscos"], sum_mode
    sp = orc.OrcScalarParams()
    s = I["sc"]
    sp.solver, sp.maxiter, sp.cscheme, sp.grad_method, sp.limiter, sp.tscheme, sp.sum_mode = orc.BICGSTAB, s["maxiter"], L.CSCHEME_ID[s["cscheme"]], 0, L.LIMITER_ID[I["limiter"]], 0, sum_mode
    sp.tol_abs, sp.tol_rel, sp.urf, sp.gds, sp.timestep, sp.viscos, sp.densit = s["tol_abs"], s["tol_rel"], s["urf"], s["gds"], 0.0, I["viscos"], I["densit"]
    return up, sp


def oracle_iteration(orc, m, c, up, sp, f, a, flomas, sum_mode):
    """One outer iteration in place; returns the momentum, pressure, k and epsilon solver reports."""
    n = m.numCells
    o = orc.calcuvw(m, c, up, f, a)
    f["apv"][:], f["apw"][:] = o["apv"], o["apw"]
    dP = o["dPdxi"]
    su = np.zeros(n)
    orc.assemble_pcorr_into(m, c, f["den"], f["u"], f["v"], f["w"], f["p"], f["pp"], dP, f["apu"], a, su, f["flmass"], flomas=flomas)
    f["pp"][:] = 0.0
    prep = orc.solve(orc.ICCG, c.ia, c.ja, a, c.diag, f["pp"], su, I["maxiterP"], I["tolAbsP"], I["tolRelP"], sum_mode)
    orc.correct_simple(m, c, 2, a, f["den"], f["u"], f["v"], f["w"], f["p"], f["pp"], f["apu"], f["apv"], f["apw"], I["urfP"], I["pRefCell"], dP, f["flmass"])
    gU, gV, gW = orc.grad_gauss(m, f["u"]), orc.grad_gauss(m, f["v"]), orc.grad_gauss(m, f["w"])        # modify_viscosity_turbulence.f90:28-33
    f["magStrain"], _ = orc.calc_strain_and_vorticity(m, gU, gV, gW)
    sp.kind, sp.prtr = orc.SC_TKE_RLZB, 1.0
    ok = orc.calcsc(m, c, sp, f)
    sp.kind, sp.prtr = orc.SC_EPS_RLZB, 1.0 / 1.2
    oe = orc.calcsc(m, c, sp, f)
    a[:] = oe["a"]        # `a` is ONE module array in the reference: the next calcuvw starts from the epsilon matrix (its stale diagonal enters velocity.f90:606)
    orc.modify_mu_eff_rlzb(m, I["urfVis"], I["viscos"], gU, gV, gW, f["te"], f["ed"], f["den"], f["u"], f["v"], f["w"], f["dnw"], f["vis"], f["visw"])
    return o["reps"], prep, ok["rep"], oe["rep"]

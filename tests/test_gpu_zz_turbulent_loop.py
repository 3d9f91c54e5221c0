"""GPU (B200): whole turbulent SIMPLE outer iterations with every field resident on the device, in the order of
src/cappuccino/main.f90:142-175 -- calcuvw, calcp_simple, then modify_viscosity_turbulence (grad_gauss of the corrected velocity,
calc_strain_and_vorticity, calcsc_tke, calcsc_epsilon, modify_mu_eff of the realizable k-epsilon model) -- on a channel with an
inlet, an outlet, two walls and symmetry sides (the ingredients of examples/pitzDaily: muscl, Venkatakrishnan, BiCGStab-ILU(0) for
U/k/epsilon, IC(0)-CG for p'), against the same chain through the oracle.  Every solve runs a fixed number of iterations (tolRel tiny)
so that the counts cannot depend on last-bit differences; k**1.5, acos, cos and log go through the device libm, hence a tolerance."""
import numpy as np
import pytest

import cases
from fcb200 import lib as L
from fcb200 import mesh as M
import test_gpu_scalar as T

pytestmark = pytest.mark.gpu
VISCOS, DENSIT = 1e-3, 1.0


def initial_state(m, orc):
    g = T.scalar_inputs(m, orc)
    n, nT, Fi = m.numCells, m.numTotal, m.numInnerFaces
    f = dict(u=np.full(nT, 1.0), v=np.zeros(nT), w=np.zeros(nT), p=np.zeros(nT), pp=np.zeros(nT), den=np.full(nT, DENSIT), vis=np.full(nT, 5 * VISCOS),
             apu=np.zeros(nT), apv=np.zeros(nT), apw=np.zeros(nT), visw=np.full(m.numBoundaryFaces, VISCOS), dnw=g["dnw"],
             te=np.full(nT, 1.5e-3), ed=np.full(nT, 2e-3), flmass=np.zeros(m.numFaces))
    for ib in range(m.numBoundaries):
        pf = m.patch_faces(ib); sl = n + pf - Fi
        if m.bctype[ib] == M.BC_WALL:
            f["u"][sl] = 0.0
        if m.bctype[ib] == M.BC_INLET:
            f["flmass"][pf] = DENSIT * (f["u"][sl] * m.arx[pf] + f["v"][sl] * m.ary[pf] + f["w"][sl] * m.arz[pf])
    flomas = float(-f["flmass"].sum())
    return f, flomas


def test_turbulent_simple_iterations_match_the_oracle(fcp, orc):
    m = cases.meshes()["channel_inout"]
    c = orc.Csr(m)
    f, flomas = initial_state(m, orc)
    n = m.numCells
    ctx = L.Context(m)
    for k in ("u", "v", "w", "p", "pp", "den", "vis", "apu", "apv", "apw", "te", "ed"):
        ctx.upload(k.upper(), f[k])
    ctx.upload("VISW", T.bslot(m, f["visw"])); ctx.upload("DNW", T.bslot(m, f["dnw"])); ctx.upload("FLMASS", f["flmass"]); ctx.upload("A", np.zeros(ctx.nnz))
    up = orc.OrcUvwParams()
    up.solver, up.maxiter, up.tol_abs, up.tol_rel = orc.BICGSTAB, 3, 1e-30, 1e-30
    up.urf[0] = up.urf[1] = up.urf[2] = 0.7
    up.gds, up.cscheme, up.limiter, up.pscheme, up.viscos, up.sum_mode = 1.0, L.CSCHEME_ID["muscl"], L.LIMITER_ID["Venkatakrishnan"], 2, VISCOS, orc.SUM_TREE
    a = np.zeros(c.nnz)
    dP = np.zeros((m.numTotal, 3))
    for it in range(4):
        # ---- device
        ctx.calcuvw(solver="bicgstab", maxiter=3, tol_abs=1e-30, tol_rel=1e-30, urf=(0.7, 0.7, 0.7), gds=1.0, cscheme="muscl", limiter="Venkatakrishnan",
                    pscheme="weighted", viscos=VISCOS)
        ctx.calcp_simple(solver="iccg", maxiter=6, tol_abs=1e-30, tol_rel=1e-30, urfp=0.3, npcor=1, pRefCell=1, pscheme="weighted", flomas=flomas)
        for comp, gfield in (("U", "DUDXI"), ("V", "DVDXI"), ("W", "DWDXI")):
            ctx.grad(L.GRAD_GAUSS, comp, gfield)
        ctx.calc_strain_and_vorticity()
        ctx.calcsc("TE", kind="tke_rlzb", solver="bicgstab", maxiter=4, tol_abs=1e-30, tol_rel=1e-30, urf=0.7, gds=1.0, cscheme="muscl",
                   limiter="Venkatakrishnan", prtr=1.0, viscos=VISCOS, densit=DENSIT)
        ctx.calcsc("ED", kind="eps_rlzb", solver="bicgstab", maxiter=4, tol_abs=1e-30, tol_rel=1e-30, urf=0.7, gds=1.0, cscheme="muscl",
                   limiter="Venkatakrishnan", prtr=1.0 / 1.2, viscos=VISCOS, densit=DENSIT)
        ctx.modify_mu_eff_k_epsilon_rlzb(0.8, VISCOS)
        # ---- oracle
        o = orc.calcuvw(m, c, up, f, a)
        f["apv"][:], f["apw"][:] = o["apv"], o["apw"]
        dP[...] = o["dPdxi"]
        su = np.zeros(n)
        orc.assemble_pcorr_into(m, c, f["den"], f["u"], f["v"], f["w"], f["p"], f["pp"], dP, f["apu"], a, su, f["flmass"], flomas=flomas)
        orc.solve(orc.ICCG, c.ia, c.ja, a, c.diag, f["pp"], su, 6, 1e-30, 1e-30, orc.SUM_TREE)
        orc.correct_simple(m, c, 2, a, f["den"], f["u"], f["v"], f["w"], f["p"], f["pp"], f["apu"], f["apv"], f["apw"], 0.3, 1, dP, f["flmass"])
        gU, gV, gW = orc.grad_gauss(m, f["u"]), orc.grad_gauss(m, f["v"]), orc.grad_gauss(m, f["w"])
        f["magStrain"], _ = orc.calc_strain_and_vorticity(m, gU, gV, gW)
        sp = T.oracle_params(orc, orc.SC_TKE_RLZB, "bicgstab", "muscl", "gauss", "Venkatakrishnan", "steady")
        sp.maxiter, sp.tol_rel, sp.urf, sp.gds, sp.prtr, sp.viscos, sp.densit = 4, 1e-30, 0.7, 1.0, 1.0, VISCOS, DENSIT
        orc.calcsc(m, c, sp, f)
        sp.kind, sp.prtr = orc.SC_EPS_RLZB, 1.0 / 1.2
        a[:] = orc.calcsc(m, c, sp, f)["a"]     # `a` is one module array in the reference: calcuvw's first row sum sees the epsilon matrix's diagonal
        orc.modify_mu_eff_rlzb(m, 0.8, VISCOS, gU, gV, gW, f["te"], f["ed"], f["den"], f["u"], f["v"], f["w"], f["dnw"], f["vis"], f["visw"])
        # from the third iteration on the stale epsilon-matrix diagonal in calcuvw's first row sum (velocity.f90:606, quirk Q25) amplifies
        # last-bit libm differences: see tests/test_gpu_zz_pitz_daily.py
        for k in ("u", "v", "w", "p", "te", "ed", "vis"):
            T.close(ctx.download(k.upper()), f[k], f"iteration {it}: {k}", 1e-7 if it < 2 else 5e-3)
        assert np.all(np.isfinite(f["te"])) and f["te"][:n].min() > 0 and f["ed"][:n].min() > 0
    ctx.close()

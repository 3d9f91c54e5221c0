"""GPU (B200): BASELINE config 2 as a parity case -- the reference's own examples/pitzDaily mesh and settings (tests/pitz_loop.py): SIMPLE
outer iterations (calcuvw, calcp_simple, grad_gauss of the corrected velocity, strain, k, epsilon, mu_eff) with every field resident on the
device against the same chain through the oracle.

Two phases.  (1) Four iterations chained freely on both sides.  The first: identical solver iteration counts, fields to 1e-10 (k**1.5, acos,
cos, log go through the device libm; everything else is bit-identical).  From the SECOND iteration on only 5e-3 (vis: 5e-2): the reference's
first momentum row sum, `sum(a(ia:ia+1-1)) - a(diag)` (velocity.f90:606), still holds the PREVIOUS equation's diagonal -- here the epsilon
matrix's, five to eleven orders larger than the momentum coefficients -- so the last bits of that stale diagonal (which depend on the libm
through the wall-cell epsilon of the iteration before) are amplified to ~1e-5 of the U-equation diagonal.  That ill-conditioning is the
reference's (quirk Q25 in DESIGN.md), reproduced faithfully; it bounds how well ANY two libm implementations can agree on this case.  The
tolerances were set with the emulation's FCP_EMU_LIBM_ULP=1/2/4 switch (transcendental results moved by up to k ulp): 1e-4 typical, 1e-2 worst
(vis).  (2) Three more iterations, the oracle re-started from the device's state before each: every iteration is then a single-iteration
comparison from identical inputs (stale diagonal included) and must again agree to 1e-10 with equal iteration counts."""
import numpy as np
import pytest

import pitz_loop as P
import test_gpu_scalar as T
from fcb200 import lib as L

pytestmark = pytest.mark.gpu


def device_iteration(ctx, I, S, flomas):
    ur = ctx.calcuvw(solver="bicgstab", maxiter=I["maxiterU"], tol_abs=I["tolAbsU"], tol_rel=I["tolRelU"], urf=I["urfU"], gds=I["gdsU"],
                     cscheme=I["cSchemeU"], limiter=I["limiter"], pscheme=I["pscheme"], viscos=I["viscos"])
    pr = ctx.calcp_simple(solver="iccg", maxiter=I["maxiterP"], tol_abs=I["tolAbsP"], tol_rel=I["tolRelP"], urfp=I["urfP"], npcor=1,
                          pRefCell=I["pRefCell"], pscheme=I["pscheme"], flomas=flomas, zero_pp=True)[0]
    for comp, gfield in (("U", "DUDXI"), ("V", "DVDXI"), ("W", "DWDXI")):
        ctx.grad(L.GRAD_GAUSS, comp, gfield)
    ctx.calc_strain_and_vorticity()
    kr, _, _ = ctx.calcsc("TE", kind="tke_rlzb", solver="bicgstab", maxiter=S["maxiter"], tol_abs=S["tol_abs"], tol_rel=S["tol_rel"], urf=S["urf"],
                          gds=S["gds"], cscheme=S["cscheme"], limiter=I["limiter"], prtr=1.0, viscos=I["viscos"], densit=I["densit"])
    er, _, _ = ctx.calcsc("ED", kind="eps_rlzb", solver="bicgstab", maxiter=S["maxiter"], tol_abs=S["tol_abs"], tol_rel=S["tol_rel"], urf=S["urf"],
                          gds=S["gds"], cscheme=S["cscheme"], limiter=I["limiter"], prtr=1.0 / 1.2, viscos=I["viscos"], densit=I["densit"])
    ctx.modify_mu_eff_k_epsilon_rlzb(I["urfVis"], I["viscos"])
    return ur, pr, kr, er


def test_pitz_daily_iterations_match_the_oracle(fcp, orc):
    m = P.mesh()
    c = orc.Csr(m)
    f, flomas = P.initial_state(m, orc)
    I, S = P.I, P.I["sc"]
    ctx = L.Context(m)
    for k in ("u", "v", "w", "p", "pp", "den", "vis", "apu", "apv", "apw", "te", "ed"):
        ctx.upload(k.upper(), f[k])
    ctx.upload("VISW", T.bslot(m, f["visw"])); ctx.upload("DNW", T.bslot(m, f["dnw"])); ctx.upload("FLMASS", f["flmass"]); ctx.upload("A", np.zeros(ctx.nnz))
    up, sp = P.oracle_params(orc, orc.SUM_TREE)
    a = np.zeros(c.nnz)
    for it in range(4):
        ur, pr, kr, er = device_iteration(ctx, I, S, flomas)
        our, opr, okr, oer = P.oracle_iteration(orc, m, c, up, sp, f, a, flomas, orc.SUM_TREE)
        if it == 0:
            assert [r.iters for r in ur] == [r.iters for r in our] and pr.iters == opr.iters and kr.iters == okr.iters and er.iters == oer.iters, \
                (it, [r.iters for r in ur], [r.iters for r in our], pr.iters, opr.iters, kr.iters, okr.iters, er.iters, oer.iters)
        for k in ("u", "v", "w", "p", "te", "ed", "vis", "flmass"):
            T.close(ctx.download(k.upper()), f[k], f"iteration {it}: {k}", 1e-10 if it == 0 else 5e-2 if k == "vis" else 5e-3)
    n = m.numCells
    for it in range(4, 7):
        for k in ("u", "v", "w", "p", "pp", "vis", "apu", "apv", "apw", "te", "ed", "flmass"):     # the oracle continues from the device's state
            f[k][:] = ctx.download(k.upper())
        f["visw"][:] = ctx.download("VISW")[n:]
        a[:] = ctx.download("A")
        ur, pr, kr, er = device_iteration(ctx, I, S, flomas)
        our, opr, okr, oer = P.oracle_iteration(orc, m, c, up, sp, f, a, flomas, orc.SUM_TREE)
        assert [r.iters for r in ur] == [r.iters for r in our] and pr.iters == opr.iters and kr.iters == okr.iters and abs(er.iters - oer.iters) <= 1, \
            (it, [r.iters for r in ur], [r.iters for r in our], pr.iters, opr.iters, kr.iters, okr.iters, er.iters, oer.iters)
        for k in ("u", "v", "w", "p", "te", "ed", "vis", "flmass"):
            T.close(ctx.download(k.upper()), f[k], f"re-synchronised iteration {it}: {k}", 1e-10)
    ctx.close()

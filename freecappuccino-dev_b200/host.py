"""Python mirror of the reference's module-level API for the hot path (same names, argument meaning and error
behaviour as the Fortran procedures), on top of the C-ABI binding in ``lib.py``.

The reference keeps its state in Fortran module globals and its procedures take few or no arguments
(``calcp_simple`` has none, Pressure/pressure.f90:39-40).  ``Case`` plays the role of those modules: it owns the host
arrays ``u, v, w, p, pp, den, apu, apv, apw, su, sv, sw, dPdxi, flmass, a`` (names of variables.f90 /
sparse_matrix.f90) and the procedures below read and write them exactly like the Fortran ones do.  All arithmetic runs
on the GPU; there is no CPU fallback.
"""
from __future__ import annotations

import sys
from typing import Optional

import numpy as np

from . import lib as L


class FvEquation:
    """type(fvEquation) of fvImplicit/fvEquation.f90:16-67: coef(nnz) over the case's CSR pattern + source(numCells), with
    operator(+), operator(-) and operator(==) producing NEW equations (:158-404).  The element-wise arithmetic runs on
    the GPU (fcp_field_axpby) in the matrix slots A and H and the scalar slots S0/S1."""

    def __init__(self, case: "Case", coef: Optional[np.ndarray] = None, source: Optional[np.ndarray] = None):
        self.case = case
        self.coef = np.zeros(case.nnz) if coef is None else np.asarray(coef, dtype=np.float64)
        self.source = np.zeros(case.mesh.numCells) if source is None else np.asarray(source, dtype=np.float64)

    def _combine(self, other, beta: float) -> "FvEquation":
        c, n = self.case.ctx, self.case.mesh.numCells
        out = FvEquation(self.case)
        if isinstance(other, FvEquation):                     # add_fvEquations / subtract_fvEquations
            c.upload("A", self.coef); c.upload("H", other.coef)
            c.axpby("A", 1.0, "A", beta, "H")
            out.coef = c.download("A")
            rhs = other.source
        else:                                                 # add_source_to_fvEquation / subtract_source_from_fvEquation:
            out.coef = np.zeros(self.case.nnz)                # the reference returns a fresh equation: coef is NOT carried over
            rhs = np.asarray(other, dtype=np.float64)[:n]
        s0 = np.zeros(self.case.mesh.numTotal); s1 = np.zeros(self.case.mesh.numTotal)
        s0[:n] = self.source; s1[:n] = rhs
        c.upload("S0", s0); c.upload("S1", s1)
        c.axpby("S0", 1.0, "S0", beta, "S1")
        out.source = c.download("S0", n)
        return out

    def __add__(self, other):
        return self._combine(other, 1.0)

    def __sub__(self, other):
        return self._combine(other, -1.0)

    def equals(self, other):
        """operator(==) is bound to the same procedures as operator(-) (fvEquation.f90:70-75)."""
        return self._combine(other, -1.0)


class Case:
    """geometry + sparse_matrix + variables modules of one run (one mesh partition on one GPU)."""

    def __init__(self, mesh, device: int = 0, out=sys.stdout):
        self.mesh = mesh
        self.out = out                      # unit 6: the solver report lines go here (main.f90:100)
        self.ctx = L.Context(mesh, device)  # create_CSR_matrix happens inside (sparse_matrix.f90:86)
        self.ia, self.ja, self.diag, self.icell_jcell_csr_index, self.jcell_icell_csr_index = self.ctx.csr_pattern()
        self.nnz = self.ctx.nnz
        nT, n = mesh.numTotal, mesh.numCells
        z = lambda k: np.zeros(k)  # noqa: E731
        self.u, self.v, self.w, self.p, self.pp = z(nT), z(nT), z(nT), z(nT), z(nT)
        self.den = np.ones(nT)
        self.apu, self.apv, self.apw = z(nT), z(nT), z(nT)
        self.su, self.sv, self.sw = z(n), z(n), z(n)
        self.dPdxi = np.zeros((nT, 3))
        self.flmass = z(mesh.numFaces)
        self.a = z(self.nnz)
        # parameters (parameters.f90, Pressure/pressure.f90:28-33, nablap.f90)
        self.pRefCell, self.npcor, self.const_mflux, self.flomas = 1, 1, False, 0.0
        self.urfP, self.lSolverP, self.maxiterP, self.tolAbsP, self.tolRelP = 0.2, "iccg", 30, 1e-13, 0.025
        self.pscheme = "linear"
        self.lstsq = self.lstsq_dm = self.lstsq_qr = False
        # PISO (parameters.f90: ncorr, npcor) and the momentum equation leftovers calcp_piso reads
        self.ncorr = 2
        self.rU, self.rV, self.rW = z(nT), z(nT), z(nT)
        self.h = z(self.nnz)
        # momentum equation (Velocity/velocity.f90:22-42, parameters.f90)
        self.vis = np.full(nT, 0.0)
        self.visw = z(int(sum(mesh.nfaces[ib] for ib in range(mesh.numBoundaries) if mesh.bctype[ib] == 0)))
        self.viscos, self.urfU, self.gdsU, self.cSchemeU = 0.0, [0.8, 0.8, 0.8], 1.0, "cds"
        self.lSolverU, self.maxiterU, self.tolAbsU, self.tolRelU = "bicgstab", 5, 1e-13, 0.025
        self.limiter, self.tscheme, self.timestep, self.piso, self.gradPcmf = "none", "steady", 0.0, False, 0.0
        self.dUdxi, self.dVdxi, self.dWdxi = np.zeros((nT, 3)), np.zeros((nT, 3)), np.zeros((nT, 3))
        for comp in "uvw":
            for lvl in (1, 2, 3):
                setattr(self, comp + "o" * lvl, z(nT))
        # turbulence (variables.f90: te, ed, dnw; TurbModelData.f90:25-44; parameters.f90: densit, magUbar)
        self.te, self.ed = z(nT), z(nT)
        self.dnw = z(self.visw.size)
        self.wallDistance = z(n)
        self.densit, self.magUbar, self.urfVis = 1.0, 0.0, 1.0
        self.TurbModelScalar = [dict(), dict()]            # per-scalar overrides of urf, gds, cScheme, lSolver, maxiter, tolAbs, tolRel

    def close(self):
        self.ctx.close()

    # ---- csrsolve(solver, fi, rhs, res0, itr_max, tol_abs, tol_rel, chvar)   linear_solvers.f90:40-59 ----------------
    def csrsolve(self, solver: str, fi: np.ndarray, rhs: np.ndarray, itr_max: int, tol_abs: float, tol_rel: float, chvar: str) -> float:
        if solver not in L.SOLVER_ID:
            raise L.FcpError(f'linear solver "{solver}" is not on the accelerated path')
        c = self.ctx
        c.upload("A", self.a)
        c.upload("S0", fi)
        c.upload("S1", rhs)
        rep = c.csrsolve(solver, "S0", "S1", itr_max, tol_abs, tol_rel)
        fi[: self.mesh.numCells] = c.download("S0", self.mesh.numCells)
        print(L.report_line(rep, chvar), file=self.out)
        self.last_report = rep
        return rep.resor                    # the `res0` dummy returns the normalised initial residual (:340)

    # ---- grad(phi, dPhidxi) / grad_gauss(u, dudxi)   gradients.f90:23-27, :1607 ------------------------------------------
    def create_lsq_grad_matrix(self):
        if self.lstsq:
            self.ctx.create_lsq_grad_matrix(L.GRAD_LSQ)
        if self.lstsq_dm:
            self.ctx.create_lsq_grad_matrix(L.GRAD_LSQ_DM)
        if self.lstsq_qr:
            self.ctx.create_lsq_grad_matrix(L.GRAD_LSQ_QR)

    def grad(self, phi: np.ndarray, dPhidxi: np.ndarray):
        # gradients.f90:118-138: lstsq, lstsq_qr, lstsq_dm, else gauss
        method = L.GRAD_LSQ if self.lstsq else L.GRAD_LSQ_QR if self.lstsq_qr else L.GRAD_LSQ_DM if self.lstsq_dm else L.GRAD_GAUSS
        self.ctx.upload("S0", phi)
        self.ctx.grad(method, "S0", "G0")
        dPhidxi[...] = self.ctx.download("G0")

    def grad_w_option(self, phi: np.ndarray, dPhidxi: np.ndarray, option: str, option_limiter: str):
        """grad(phi, dPhidxi, option, option_limiter)   gradients.f90:217-278; unknown option strings leave dPhidxi = 0 and
        unknown limiter strings mean 'no-limit', as in the reference's select case."""
        self.ctx.upload("S0", phi)
        if option not in L.GRAD_ID:
            dPhidxi[...] = 0.0
            return
        if option in ("lsq", "wlsq", "lsq_qr"):
            self.ctx.create_lsq_grad_matrix(L.GRAD_ID[option])
        self.ctx.grad_opt(option, L.LIMITER_ID.get(option_limiter, 0), "S0", "G0")
        dPhidxi[...] = self.ctx.download("G0")

    def grad_gauss(self, u: np.ndarray, dudxi: np.ndarray):
        self.ctx.upload("S0", u)
        self.ctx.grad(L.GRAD_GAUSS, "S0", "G0")
        dudxi[...] = self.ctx.download("G0")

    # ---- laplacian(mu, phi): fills a, accumulates into su   fvImplicit/laplacian.f90 ----------------------------------------
    def laplacian(self, mu: np.ndarray, phi: np.ndarray):
        c = self.ctx
        c.upload("S0", mu)
        c.upload("S1", phi)
        c.upload("SU", self.su)
        c.laplacian("S0", "S1")
        self.a[...] = c.download("A")
        self.su[...] = c.download("SU", self.mesh.numCells)

    # ---- gradp_and_sources(p)   Pressure/nablap.f90:19 ------------------------------------------------------------------------
    def gradp_and_sources(self, p: np.ndarray):
        if self.pscheme not in L.PSCHEME:   # nablap.f90:106-110 stops
            raise L.FcpError(f"unknown pscheme {self.pscheme}")
        c = self.ctx
        n = self.mesh.numCells
        c.upload("P", p)
        c.upload("APU", self.apu)
        c.upload("DPDXI", self.dPdxi)
        c.gradp_and_sources(self.pscheme, "P")
        p[...] = c.download("P")
        self.su[...], self.sv[...], self.sw[...] = c.download("SU", n), c.download("SV", n), c.download("SW", n)
        self.dPdxi[...] = c.download("DPDXI")

    # ---- calcp_simple()   Pressure/calcp_simple.f90 ------------------------------------------------------------------------------
    def calcp_simple(self, zero_pp: bool = False):
        c = self.ctx
        n = self.mesh.numCells
        for name in ("u", "v", "w", "p", "pp", "den", "apu", "apv", "apw"):
            c.upload(name.upper(), getattr(self, name))
        c.upload("DPDXI", self.dPdxi)
        c.upload("FLMASS", self.flmass)
        reps = c.calcp_simple(solver=self.lSolverP, maxiter=self.maxiterP, tol_abs=self.tolAbsP, tol_rel=self.tolRelP, urfp=self.urfP,
                              npcor=self.npcor, pRefCell=self.pRefCell, pscheme=self.pscheme, const_mflux=self.const_mflux,
                              flomas=self.flomas, zero_pp=zero_pp)
        for r in reps:
            print(L.report_line(r, "p"), file=self.out)
        for name in ("u", "v", "w", "p", "pp"):
            getattr(self, name)[...] = c.download(name.upper())
        self.flmass[...] = c.download("FLMASS")
        self.dPdxi[...] = c.download("DPDXI")
        self.su[...], self.sv[...], self.sw[...] = c.download("SU", n), c.download("SV", n), c.download("SW", n)
        self.a[...] = c.download("A")
        return reps

    # ---- calcuvw()   Velocity/velocity.f90:50-750 -------------------------------------------------------------------------------
    def calcuvw(self):
        """The momentum predictor.  Reads the module state (u,v,w,p,den,vis,visw,flmass and the past time levels), leaves
        u,v,w, apu,apv,apw, su,sv,sw, dUdxi..dWdxi, dPdxi, a (and rU,rV,rW when piso) like the Fortran routine."""
        c = self.ctx
        m = self.mesh
        n = m.numCells
        for name in ("u", "v", "w", "p", "den", "vis", "apu"):
            c.upload(name.upper(), getattr(self, name))
        visw = np.zeros(m.numTotal)             # visw(iWall), wall faces in patch order (velocity.f90:441-443) -> their boundary slots
        iw = 0
        for ib in range(m.numBoundaries):
            if m.bctype[ib] == 0:               # wall
                sl = n + m.patch_faces(ib) - m.numInnerFaces
                visw[sl] = self.visw[iw: iw + sl.size]
                iw += sl.size
        c.upload("VISW", visw)
        c.upload("FLMASS", self.flmass)
        c.upload("A", self.a)
        nlev = L.TSCHEME[self.tscheme]
        for lvl in range(nlev):
            for comp in "uvw":
                key = comp + "o" * (lvl + 1)
                c.upload(key.upper(), getattr(self, key))
        grad = "lsq" if self.lstsq else "lsq_qr" if self.lstsq_qr else "wlsq" if self.lstsq_dm else "gauss"
        reps = c.calcuvw(solver=self.lSolverU, maxiter=self.maxiterU, tol_abs=self.tolAbsU, tol_rel=self.tolRelU, urf=tuple(self.urfU), gds=self.gdsU,
                         cscheme=self.cSchemeU, grad_method=grad, limiter=L.LIMITER_ID.get(self.limiter, 0), pscheme=self.pscheme, tscheme=self.tscheme,
                         timestep=self.timestep, piso=self.piso, const_mflux=self.const_mflux, gradPcmf=self.gradPcmf, viscos=self.viscos)
        for r, ch in zip(reps, "UVW"):
            print(L.report_line(r, ch), file=self.out)
        for name in ("u", "v", "w", "p", "apu", "apv", "apw"):
            getattr(self, name)[...] = c.download(name.upper())
        self.su[...], self.sv[...], self.sw[...] = c.download("SU", n), c.download("SV", n), c.download("SW", n)
        self.dUdxi[...], self.dVdxi[...], self.dWdxi[...] = c.download("DUDXI"), c.download("DVDXI"), c.download("DWDXI")
        self.dPdxi[...] = c.download("DPDXI")
        self.a[...] = c.download("A")
        if self.piso:
            self.rU[:n], self.rV[:n], self.rW[:n] = c.download("RU", n), c.download("RV", n), c.download("RW", n)
        return reps

    # ---- calcp_piso()   Pressure/calcp_piso.f90 ----------------------------------------------------------------------------------
    def calcp_piso(self):
        c = self.ctx
        n = self.mesh.numCells
        for name in ("u", "v", "w", "p", "pp", "den", "apu", "apv", "apw"):
            c.upload(name.upper(), getattr(self, name))
        c.upload("RU", self.rU); c.upload("RV", self.rV); c.upload("RW", self.rW)
        c.upload("A", self.a)               # the momentum coefficients (h = a, calcp_piso.f90:81)
        c.upload("DPDXI", self.dPdxi)
        c.upload("FLMASS", self.flmass)
        reps = c.calcp_piso(solver=self.lSolverP, maxiter=self.maxiterP, tol_abs=self.tolAbsP, tol_rel=self.tolRelP, urfp=self.urfP,
                            ncorr=self.ncorr, npcor=self.npcor, pscheme=self.pscheme, const_mflux=self.const_mflux, flomas=self.flomas)
        for r in reps:
            print(L.report_line(r, "p"), file=self.out)
        for name in ("u", "v", "w", "p", "pp"):
            getattr(self, name)[...] = c.download(name.upper())
        self.flmass[...] = c.download("FLMASS")
        self.dPdxi[...] = c.download("DPDXI")
        self.su[...], self.sv[...], self.sw[...] = c.download("SU", n), c.download("SV", n), c.download("SW", n)
        self.a[...] = c.download("A")
        self.h[...] = c.download("H")
        return reps

    # ---- row f3 / f4 procedures ----------------------------------------------------------------------------------------------
    def _wall_slots(self, per_wall_face: np.ndarray) -> np.ndarray:
        """x(iWall), wall faces counted in patch order (the reference's convention) -> a numTotal field with the values in the wall faces' boundary slots."""
        m, n = self.mesh, self.mesh.numCells
        out = np.zeros(m.numTotal)
        iw = 0
        for ib in range(m.numBoundaries):
            if m.bctype[ib] == 0:
                sl = n + m.patch_faces(ib) - m.numInnerFaces
                out[sl] = per_wall_face[iw: iw + sl.size]
                iw += sl.size
        return out

    def _from_wall_slots(self, field: np.ndarray, per_wall_face: np.ndarray):
        m, n = self.mesh, self.mesh.numCells
        iw = 0
        for ib in range(m.numBoundaries):
            if m.bctype[ib] == 0:
                sl = n + m.patch_faces(ib) - m.numInnerFaces
                per_wall_face[iw: iw + sl.size] = field[sl]
                iw += sl.size

    def updateBoundary(self, phi: np.ndarray):
        """boundary/updateBoundary.f90"""
        self.ctx.upload("S0", phi)
        self.ctx.update_boundary("S0")
        phi[...] = self.ctx.download("S0")

    def wall_distance(self) -> np.ndarray:
        """mesh/wall_distance.f90:55-133 -> wallDistance(numCells); prints the solver's report line like the reference."""
        rep = self.ctx.wall_distance()
        print(L.report_line(rep, "Wdis"), file=self.out)
        self.wallDistance = self.ctx.download("WALLDIST", self.mesh.numCells)
        return self.wallDistance

    def constant_mass_flow_forcing(self):
        """cappuccino/constant_mass_flow_forcing.f90: corrects u, updates gradPcmf, prints the reference's line."""
        self.ctx.upload("U", self.u); self.ctx.upload("APU", self.apu)
        self.gradPcmf, ustar = self.ctx.constant_mass_flow_forcing(self.magUbar, self.gradPcmf)
        self.u[...] = self.ctx.download("U")
        print(f"  Uncorrected Ubar = {ustar:13.6E} pressure gradient = {self.gradPcmf:13.6E}", file=self.out)

    def _upload_turbulence_state(self):
        c = self.ctx
        for name in ("u", "v", "w", "den", "vis", "te", "ed"):
            c.upload(name.upper(), getattr(self, name))
        c.upload("FLMASS", self.flmass)
        c.upload("VISW", self._wall_slots(self.visw)); c.upload("DNW", self._wall_slots(self.dnw))
        for comp, gfield in (("U", "DUDXI"), ("V", "DVDXI"), ("W", "DWDXI")):       # modify_viscosity_turbulence.f90:28-33
            c.grad(L.GRAD_GAUSS, comp, gfield)
        c.calc_strain_and_vorticity()

    def _scalar_settings(self, i: int) -> dict:
        s = self.TurbModelScalar[i]                                                  # TurbModel%Scalar(i), TurbModelData.f90:25-33
        grad = "lsq" if self.lstsq else "lsq_qr" if self.lstsq_qr else "wlsq" if self.lstsq_dm else "gauss"
        return dict(solver=s.get("lSolver", "bicgstab"), maxiter=s.get("maxiter", 10), tol_abs=s.get("tolAbs", 1e-10), tol_rel=s.get("tolRel", 0.01),
                    urf=s.get("urf", 0.7), gds=s.get("gds", 1.0), cscheme=s.get("cScheme", "linearUpwind"), grad_method=grad,
                    limiter=L.LIMITER_ID.get(self.limiter, 0), tscheme=self.tscheme if self.tscheme in ("steady", "bdf", "bdf2") else "steady",
                    timestep=self.timestep, viscos=self.viscos, densit=self.densit)

    def _download_turbulence_state(self):
        c = self.ctx
        self.te[...], self.ed[...], self.vis[...] = c.download("TE"), c.download("ED"), c.download("VIS")
        self._from_wall_slots(c.download("VISW"), self.visw)

    def modify_viscosity_k_epsilon_rlzb(self):
        """TurbulenceModels/k_epsilon_rlzb.f90:38-50: calcsc_tke, calcsc_epsilon, modify_mu_eff."""
        c = self.ctx
        self._upload_turbulence_state()
        for i, (field, kind, prtr, nm) in enumerate((("TE", "tke_rlzb", 1.0, "k"), ("ED", "eps_rlzb", 1.0 / 1.2, "epsilon"))):
            rep, lo, hi = c.calcsc(field, kind=kind, prtr=prtr, **self._scalar_settings(i))
            print(L.report_line(rep, nm), file=self.out)
            print(f"  {lo:11.4E} <= {nm} <= {hi:11.4E}", file=self.out)
        c.modify_mu_eff_k_epsilon_rlzb(self.urfVis, self.viscos)
        self._download_turbulence_state()

    def modify_viscosity_k_omega_sst(self, LowRe: bool = False):
        """TurbulenceModels/k_omega_SST.f90:62-88 (needs wall_distance() once before)."""
        c = self.ctx
        self._upload_turbulence_state()
        wd = np.zeros(self.mesh.numTotal); wd[: self.mesh.numCells] = self.wallDistance
        c.upload("WALLDIST", wd)
        for i, (field, kind, nm) in enumerate((("TE", "tke_sst", "k"), ("ED", "omega_sst", "Omega"))):
            rep, lo, hi = c.calcsc(field, kind=kind, lowre=LowRe, **self._scalar_settings(i))
            print(L.report_line(rep, nm), file=self.out)
            print(f"  {lo:11.4E} <= {nm} <= {hi:11.4E}", file=self.out)
        c.modify_mu_eff_k_omega_sst(self.urfVis, self.viscos, self.densit, LowRe)
        self._download_turbulence_state()

    def modify_viscosity_sgs(self, model: str):
        """TurbulenceModels/wale_sgs.f90:33 ('wale') / vremanSGS.f90:33 ('vreman')."""
        c = self.ctx
        for name in ("u", "v", "w", "den", "vis"):
            c.upload(name.upper(), getattr(self, name))
        c.modify_viscosity_sgs(model, self.urfVis, self.viscos)
        self.vis[...] = c.download("VIS")
        self._from_wall_slots(c.download("VISW"), self.visw)
        print(f"  {(self.vis / self.viscos).min():11.4E} <= Viscosity ratio <= {(self.vis / self.viscos).max():11.4E}", file=self.out)

    # ---- src-par: exchange(phi), global_sum(x)   src-par/exchange.f90:3, global_sum_mpi.f90:4 ---------------------------------------
    def comm_init(self, rank: int, nranks: int, unique_id: Optional[bytes]):
        self.ctx.comm_init(rank, nranks, unique_id, self.mesh.peer_rank)

    def exchange(self, phi: np.ndarray):
        self.ctx.upload("S0", phi)
        self.ctx.exchange("S0")
        phi[...] = self.ctx.download("S0")

    def global_sum(self, x: float) -> float:
        return self.ctx.global_sum(x)

    def global_isum(self, i: int) -> int:          # src-par/global_isum_mpi.f90
        return self.ctx.global_isum(i)

    def global_max(self, x: float) -> float:       # src-par/global_max_mpi.f90
        return self.ctx.global_max(x)

    def global_min(self, x: float) -> float:       # src-par/global_min_mpi.f90
        return self.ctx.global_min(x)

"""ctypes binding of the C-ABI shared library ``csrc/libfcp_b200.so`` (include/fcp.h).

This is the only way the Python host layer reaches the GPU; there is no CPU fallback.  Loading fails loudly when
the library is missing (``build()`` compiles it with nvcc for sm_100a), and every compute call fails with
``FcpError`` when no B200 is visible.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
SO_PATH = os.path.join(CSRC, "libfcp_b200.so")
HEADER = os.path.join(os.path.dirname(_HERE), "include", "fcp.h")

_pd = C.POINTER(C.c_double)
_pi = C.POINTER(C.c_int32)

# ---- constants of include/fcp.h -------------------------------------------------------------------------------
OK, EINVAL, ECUDA, ENODEVICE, ENCCL, ESTATE = 0, -1, -2, -3, -4, -5
SOLVER_DPCG, SOLVER_ICCG, SOLVER_BICGSTAB, SOLVER_GAUSS_SEIDEL = 1, 2, 3, 4
SOLVER_ID = {"dpcg": SOLVER_DPCG, "iccg": SOLVER_ICCG, "bicgstab": SOLVER_BICGSTAB, "gauss-seidel": SOLVER_GAUSS_SEIDEL}
GRAD_GAUSS, GRAD_LSQ, GRAD_LSQ_DM, GRAD_LSQ_QR = 0, 1, 2, 3
GRAD_ID = {"gauss": GRAD_GAUSS, "lsq": GRAD_LSQ, "wlsq": GRAD_LSQ_DM, "lsq_qr": GRAD_LSQ_QR}          # option strings, gradients.f90:240-256
LIMITER_NONE, LIMITER_BJ, LIMITER_VENKAT, LIMITER_R3, LIMITER_MDL = 0, 1, 2, 3, 4
LIMITER_ID = {"none": 0, "no-limit": 0, "Barth-Jespersen": 1, "Venkatakrishnan": 2, "R3": 3, "multidimensional": 4}   # gradients.f90:261-276
PSCHEME = {"linear": 0, "central": 1, "weighted": 2}
FIELDS = ["U", "V", "W", "P", "PP", "DEN", "VIS", "APU", "APV", "APW", "SU", "SV", "SW", "S0", "S1", "S2", "S3",
          "DUDXI", "DVDXI", "DWDXI", "DPDXI", "G0", "G1", "FLMASS", "A", "APR", "H", "RU", "RV", "RW", "VISW",
          "UO", "VO", "WO", "UOO", "VOO", "WOO", "UOOO", "VOOO", "WOOO", "SPU", "SPV", "SP",
          "TE", "ED", "PHIO", "PHIOO", "GEN", "MAGSTRAIN", "VORTICITY", "DNW", "TAU", "YPL", "SCTMP",
          "FSST", "WALLDIST", "DTEDXI", "DEDDXI"]
F = {name: i for i, name in enumerate(FIELDS)}
KERNEL_CLASSES = ["spmv_dot", "cg_pk", "cg_update", "cg_init", "precond", "dot", "bicg_elem", "assemble", "gradp", "correct_flux",
                  "grad", "laplacian", "spmv", "halo", "limiter", "piso_h", "uvw", "scalar", "krylov_persist"]
GRADIENT_FIELDS = {"DUDXI", "DVDXI", "DWDXI", "DPDXI", "G0", "G1", "DTEDXI", "DEDDXI"}


class FcpError(RuntimeError):
    pass


class MeshDesc(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("numCells", "numInnerFaces", "numBoundaryFaces", "numBoundaries")] + \
               [("owner", _pi), ("neighbour", _pi)] + \
               [(n, _pd) for n in ("arx", "ary", "arz", "xf", "yf", "zf", "facint", "Df", "xc", "yc", "zc", "vol")] + \
               [(n, _pi) for n in ("bctype", "nfaces", "startFace", "startFaceTwin")] + [("DfPeriodic", _pd)]


class Report(C.Structure):
    _fields_ = [("res0", C.c_double), ("resl", C.c_double), ("factor", C.c_double), ("resor", C.c_double),
                ("iters", C.c_int32), ("solver", C.c_int32)]

    def as_dict(self):
        return dict(res0=self.res0, resl=self.resl, factor=self.factor, resor=self.resor, iters=self.iters, solver=self.solver)


class SimpleParams(C.Structure):
    _fields_ = [("solver", C.c_int32), ("maxiter", C.c_int32), ("tol_abs", C.c_double), ("tol_rel", C.c_double),
                ("urfp", C.c_double), ("npcor", C.c_int32), ("pRefCell", C.c_int32), ("pscheme", C.c_int32),
                ("const_mflux", C.c_int32), ("flomas", C.c_double), ("zero_pp", C.c_int32)]


CSCHEMES = ["cds", "central", "linearUpwind", "kappa", "muscl", "umist", "koren", "smart", "avl-smart", "charm", "vanleer", "ospre", "minmod",
            "boundedLinearUpwind", "boundedLinearUpwind02", "boundedCentral", "fromm", "cui", "quick", "spl13"]      # cSchemeU strings, interpolation.f90
CSCHEME_ID = {n: i for i, n in enumerate(CSCHEMES)}
TSCHEME = {"steady": 0, "bdf": 1, "bdf2": 2, "bdf3": 3}


class UvwParams(C.Structure):
    _fields_ = [("solver", C.c_int32), ("maxiter", C.c_int32), ("tol_abs", C.c_double), ("tol_rel", C.c_double), ("urf", C.c_double * 3),
                ("gds", C.c_double), ("cscheme", C.c_int32), ("grad_method", C.c_int32), ("limiter", C.c_int32), ("pscheme", C.c_int32),
                ("tscheme", C.c_int32), ("piso", C.c_int32), ("timestep", C.c_double), ("const_mflux", C.c_int32), ("pad", C.c_int32),
                ("gradPcmf", C.c_double), ("viscos", C.c_double)]


SC_GENERIC, SC_TKE_RLZB, SC_EPS_RLZB, SC_TKE_SST, SC_OMEGA_SST = 0, 1, 2, 3, 4
SC_KIND = {"generic": SC_GENERIC, "tke_rlzb": SC_TKE_RLZB, "eps_rlzb": SC_EPS_RLZB, "tke_sst": SC_TKE_SST, "omega_sst": SC_OMEGA_SST}


class ScalarParams(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("kind", "solver", "maxiter", "cscheme", "grad_method", "limiter", "tscheme", "lowre")] + \
               [(n, C.c_double) for n in ("tol_abs", "tol_rel", "urf", "gds", "timestep", "prtr", "viscos", "densit")]


class PisoParams(C.Structure):
    _fields_ = [("solver", C.c_int32), ("maxiter", C.c_int32), ("tol_abs", C.c_double), ("tol_rel", C.c_double),
                ("urfp", C.c_double), ("ncorr", C.c_int32), ("npcor", C.c_int32), ("pscheme", C.c_int32),
                ("const_mflux", C.c_int32), ("flomas", C.c_double)]


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile csrc/*.cu for sm_100a into csrc/libfcp_b200.so (nvcc cross-compiles without a GPU)."""
    srcs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh", ".h"))] + [HEADER]
    stale = not os.path.exists(SO_PATH) or any(os.path.getmtime(s) > os.path.getmtime(SO_PATH) for s in srcs)
    if force or stale:
        cmd = ["make", "-C", CSRC, "-j8"] + (["-B"] if force else []) + ["libfcp_b200.so"]
        subprocess.check_call(cmd, stdout=None if verbose else subprocess.DEVNULL)
    return SO_PATH


_LIB = None


def declared_symbols():
    """Every function name include/fcp.h declares (used by the CPU test that checks the exports)."""
    import re
    with open(HEADER) as fh:
        text = fh.read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(fcp_[a-z0-9_]+)\s*\(", text)))


def lib():
    """Load the shared library (no build here: a missing library is an error, not a reason to fall back)."""
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.exists(SO_PATH):
        raise FcpError(f"{SO_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` (nvcc, sm_100a). "
                       "There is no CPU fallback.")
    L = C.CDLL(SO_PATH)
    L.fcp_last_error.restype = C.c_char_p
    L.fcp_launch_count.restype = C.c_int64
    vp = C.c_void_p
    L.fcp_ctx_create.argtypes = [C.POINTER(MeshDesc), C.c_int, C.POINTER(vp)]
    L.fcp_ctx_destroy.argtypes = [vp]
    L.fcp_ctx_sizes.argtypes = [vp, _pi, _pi, _pi, _pi, _pi]
    L.fcp_csr_pattern.argtypes = [vp, _pi, _pi, _pi, _pi, _pi]
    L.fcp_sync.argtypes = [vp]
    L.fcp_field_upload.argtypes = [vp, C.c_int, _pd, C.c_int64]
    L.fcp_field_download.argtypes = [vp, C.c_int, _pd, C.c_int64]
    L.fcp_field_fill.argtypes = [vp, C.c_int, C.c_double]
    L.fcp_field_copy.argtypes = [vp, C.c_int, C.c_int]
    L.fcp_field_axpby.argtypes = [vp, C.c_int, C.c_double, C.c_int, C.c_double, C.c_int]
    L.fcp_field_devptr.argtypes = [vp, C.c_int, C.POINTER(vp), C.POINTER(C.c_int64)]
    L.fcp_spmv.argtypes = [vp, C.c_int, C.c_int]
    L.fcp_csrsolve.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_int32, C.c_double, C.c_double, C.POINTER(Report)]
    L.fcp_report_line.argtypes = [C.POINTER(Report), C.c_char_p, C.c_char_p, C.c_int]
    L.fcp_create_lsq_grad_matrix.argtypes = [vp, C.c_int]
    L.fcp_grad.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_int]
    L.fcp_slope_limiter.argtypes = [vp, C.c_int, C.c_int, C.c_int]
    L.fcp_grad_opt.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_int]
    L.fcp_calcp_piso.argtypes = [vp, C.POINTER(PisoParams), C.POINTER(Report)]
    L.fcp_constant_mass_flow_forcing.argtypes = [vp, C.c_double, _pd, _pd]
    L.fcp_update_boundary.argtypes = [vp, C.c_int]
    L.fcp_calcsc.argtypes = [vp, C.POINTER(ScalarParams), C.c_int, C.POINTER(Report), _pd, _pd]
    L.fcp_calc_strain_and_vorticity.argtypes = [vp]
    L.fcp_wall_distance.argtypes = [vp, C.POINTER(Report)]
    L.fcp_grad_gauss_fvx.argtypes = [vp, C.c_int, C.c_int]
    L.fcp_grad_gauss_iter.argtypes = [vp, C.c_int, C.c_int, C.c_int]
    L.fcp_modify_viscosity_sgs.argtypes = [vp, C.c_int, C.c_double, C.c_double]
    L.fcp_modify_mu_eff_k_epsilon_rlzb.argtypes = [vp, C.c_double, C.c_double]
    L.fcp_modify_mu_eff_k_omega_sst.argtypes = [vp, C.c_double, C.c_double, C.c_double, C.c_int]
    L.fcp_calcuvw.argtypes = [vp, C.POINTER(UvwParams), C.POINTER(Report)]
    L.fcp_laplacian.argtypes = [vp, C.c_int, C.c_int]
    L.fcp_gradp_and_sources.argtypes = [vp, C.c_int, C.c_int]
    L.fcp_assemble_pcorr_simple.argtypes = [vp, C.c_int, C.c_double]
    L.fcp_correct_simple.argtypes = [vp, C.c_int, C.c_double, C.c_int32]
    L.fcp_nonorth_corrector.argtypes = [vp]
    L.fcp_calcp_simple.argtypes = [vp, C.POINTER(SimpleParams), C.POINTER(Report)]
    L.fcp_solver_create.argtypes = [C.c_int32, C.c_int32, _pi, _pi, _pi, C.c_int, C.POINTER(vp)]
    L.fcp_solver_destroy.argtypes = [vp]
    L.fcp_solver_solve.argtypes = [vp, C.c_int, _pd, _pd, _pd, C.c_int32, C.c_double, C.c_double, C.POINTER(Report)]
    L.fcp_comm_unique_id.argtypes = [C.c_void_p]
    L.fcp_comm_init.argtypes = [vp, C.c_int, C.c_int, C.c_void_p, _pi]
    L.fcp_comm_mode.argtypes = [vp]
    L.fcp_exchange.argtypes = [vp, C.c_int]
    L.fcp_set_process_facint.argtypes = [vp, _pd, C.c_int32]
    L.fcp_set_process_orientation.argtypes = [vp, _pi, C.c_int32]
    L.fcp_set_flux_variant.argtypes = [vp, C.c_int, C.c_int]
    for nm in ("fcp_global_sum", "fcp_global_max", "fcp_global_min"):
        getattr(L, nm).argtypes = [vp, _pd]
    L.fcp_global_isum.argtypes = [vp, C.POINTER(C.c_int64)]
    L.fcp_profile_enable.argtypes = [vp, C.c_int]
    L.fcp_profile_reset.argtypes = [vp]
    L.fcp_profile_read.argtypes = [vp, C.c_int, _pd, C.POINTER(C.c_int64)]
    L.fcp_timer_start.argtypes = [vp]
    L.fcp_timer_stop.argtypes = [vp, C.POINTER(C.c_float)]
    L.fcp_flush_l2.argtypes = [vp]
    _LIB = L
    return L


def check(rc: int, what: str = ""):
    if rc != OK:
        msg = lib().fcp_last_error()
        raise FcpError(f"{what or 'libfcp_b200'} failed with code {rc}: {msg.decode() if msg else ''}")


def launch_count() -> int:
    return int(lib().fcp_launch_count())


def _d(a: np.ndarray):
    assert a.dtype == np.float64 and a.flags.c_contiguous, (a.dtype, a.flags.c_contiguous)
    return a.ctypes.data_as(_pd)


def _i(a: np.ndarray):
    assert a.dtype == np.int32 and a.flags.c_contiguous
    return a.ctypes.data_as(_pi)


def field_id(f) -> int:
    return F[f.upper()] if isinstance(f, str) else int(f)


class Context:
    """One mesh (partition) on one GPU: owns the ``fcp_ctx`` handle."""

    def __init__(self, mesh, device: int = 0):
        self.mesh = mesh
        self._keep = {}
        md = MeshDesc()
        md.numCells, md.numInnerFaces, md.numBoundaryFaces, md.numBoundaries = mesh.numCells, mesh.numInnerFaces, mesh.numBoundaryFaces, mesh.numBoundaries
        for n in ("owner", "neighbour", "bctype", "nfaces", "startFace"):
            self._keep[n] = np.ascontiguousarray(getattr(mesh, n), dtype=np.int32)
            setattr(md, n, _i(self._keep[n]))
        for n in ("arx", "ary", "arz", "xf", "yf", "zf", "facint", "Df", "xc", "yc", "zc", "vol"):
            self._keep[n] = np.ascontiguousarray(getattr(mesh, n), dtype=np.float64)
            setattr(md, n, _d(self._keep[n]))
        if getattr(mesh, "startFaceTwin", None) is not None:       # periodic pairs (geometry.f90:251-257)
            self._keep["startFaceTwin"] = np.ascontiguousarray(mesh.startFaceTwin, dtype=np.int32)
            md.startFaceTwin = _i(self._keep["startFaceTwin"])
        if getattr(mesh, "DfPeriodic", None) is not None:          # partitions: the global mesh's "Df(i)" per periodic face (quirk Q21)
            self._keep["DfPeriodic"] = np.ascontiguousarray(mesh.DfPeriodic, dtype=np.float64)
            md.DfPeriodic = _d(self._keep["DfPeriodic"])
        self.numPeriodic = int(getattr(mesh, "numPeriodic", 0))
        self.h = C.c_void_p()
        check(lib().fcp_ctx_create(C.byref(md), device, C.byref(self.h)), "fcp_ctx_create")
        self._keep.clear()
        s = [C.c_int32() for _ in range(5)]
        check(lib().fcp_ctx_sizes(self.h, *[C.byref(x) for x in s]))
        self.numCells, self.numTotal, self.numFaces, self.nnz, self.npro = (int(x.value) for x in s)

    def close(self):
        if getattr(self, "h", None) is not None and self.h:
            lib().fcp_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- pattern / fields ------------------------------------------------------------------------------------
    def csr_pattern(self):
        ia = np.zeros(self.numCells + 1, np.int32)
        ja = np.zeros(self.nnz, np.int32)
        diag = np.zeros(self.numCells, np.int32)
        Fi = self.mesh.numInnerFaces + self.numPeriodic      # sparse_matrix.f90:246-247
        kpn = np.zeros(Fi, np.int32)
        knp = np.zeros(Fi, np.int32)
        check(lib().fcp_csr_pattern(self.h, _i(ia), _i(ja), _i(diag), _i(kpn), _i(knp)))
        return ia, ja, diag, kpn, knp

    def extent(self, field) -> int:
        name = FIELDS[field_id(field)]
        if name in GRADIENT_FIELDS:
            return 3 * self.numTotal
        return {"FLMASS": self.numFaces, "A": self.nnz, "H": self.nnz, "APR": self.npro}.get(name, self.numTotal)

    def upload(self, field, host: np.ndarray):
        host = np.ascontiguousarray(host, dtype=np.float64).ravel()
        check(lib().fcp_field_upload(self.h, field_id(field), _d(host), host.size), f"upload({field})")

    def download(self, field, count: int | None = None) -> np.ndarray:
        n = self.extent(field) if count is None else count
        out = np.empty(n)
        check(lib().fcp_field_download(self.h, field_id(field), _d(out), n), f"download({field})")
        if FIELDS[field_id(field)] in GRADIENT_FIELDS and count is None:
            return out.reshape(-1, 3)
        return out

    def fill(self, field, value: float):
        check(lib().fcp_field_fill(self.h, field_id(field), float(value)))

    def copy(self, dst, src):
        check(lib().fcp_field_copy(self.h, field_id(dst), field_id(src)))

    def axpby(self, dst, alpha: float, x, beta: float, y):
        check(lib().fcp_field_axpby(self.h, field_id(dst), float(alpha), field_id(x), float(beta), field_id(y)), "fcp_field_axpby")

    def sync(self):
        check(lib().fcp_sync(self.h))

    # ---- operators -------------------------------------------------------------------------------------------
    def spmv(self, x, y):
        check(lib().fcp_spmv(self.h, field_id(x), field_id(y)), "fcp_spmv")

    def csrsolve(self, solver, fi, rhs, itr_max, tol_abs, tol_rel) -> Report:
        rep = Report()
        sid = SOLVER_ID[solver] if isinstance(solver, str) else int(solver)
        check(lib().fcp_csrsolve(self.h, sid, field_id(fi), field_id(rhs), itr_max, tol_abs, tol_rel, C.byref(rep)), "fcp_csrsolve")
        return rep

    def create_lsq_grad_matrix(self, method):
        check(lib().fcp_create_lsq_grad_matrix(self.h, method), "fcp_create_lsq_grad_matrix")

    def grad(self, method, phi, grad, lsq_row2_reference: bool = True):
        check(lib().fcp_grad(self.h, method, field_id(phi), field_id(grad), int(lsq_row2_reference)), "fcp_grad")

    def slope_limiter(self, limiter, phi, grad):
        lim = LIMITER_ID[limiter] if isinstance(limiter, str) else int(limiter)
        check(lib().fcp_slope_limiter(self.h, lim, field_id(phi), field_id(grad)), "fcp_slope_limiter")

    def grad_opt(self, option, option_limiter, phi, grad):
        """grad(phi, dPhidxi, option, option_limiter), gradients.f90:217-278."""
        meth = GRAD_ID[option] if isinstance(option, str) else int(option)
        lim = LIMITER_ID[option_limiter] if isinstance(option_limiter, str) else int(option_limiter)
        check(lib().fcp_grad_opt(self.h, meth, lim, field_id(phi), field_id(grad)), "fcp_grad_opt")

    def calcp_piso(self, solver="iccg", maxiter=100, tol_abs=1e-13, tol_rel=1e-6, urfp=1.0, ncorr=2, npcor=1, pscheme="linear",
                   const_mflux=False, flomas=0.0):
        prm = PisoParams(SOLVER_ID[solver] if isinstance(solver, str) else solver, maxiter, tol_abs, tol_rel, urfp, ncorr, npcor,
                         PSCHEME[pscheme] if isinstance(pscheme, str) else pscheme, int(const_mflux), flomas)
        reps = (Report * (ncorr * npcor))()
        check(lib().fcp_calcp_piso(self.h, C.byref(prm), reps), "fcp_calcp_piso")
        return [reps[i] for i in range(ncorr * npcor)]

    def constant_mass_flow_forcing(self, magUbar: float, gradPcmf: float):
        """constant_mass_flow_forcing.f90: corrects U on the device; returns (new gradPcmf, uncorrected bulk velocity)."""
        g = C.c_double(gradPcmf)
        ustar = C.c_double(0.0)
        check(lib().fcp_constant_mass_flow_forcing(self.h, magUbar, C.byref(g), C.byref(ustar)), "fcp_constant_mass_flow_forcing")
        return g.value, ustar.value

    def calcsc(self, phi, kind="generic", solver="bicgstab", maxiter=10, tol_abs=1e-13, tol_rel=0.025, urf=0.8, gds=1.0, cscheme="cds",
               grad_method="gauss", limiter="none", tscheme="steady", timestep=0.0, prtr=1.0, viscos=0.0, densit=1.0, lowre=False):
        """The calcsc template (k_epsilon_rlzb.f90:52-790 + scalar_fluxes.f90).  Returns (report, fimin, fimax)."""
        prm = ScalarParams()
        prm.kind = SC_KIND[kind] if isinstance(kind, str) else kind
        prm.solver = SOLVER_ID[solver] if isinstance(solver, str) else solver
        prm.maxiter, prm.tol_abs, prm.tol_rel, prm.urf, prm.gds = maxiter, tol_abs, tol_rel, urf, gds
        prm.cscheme = CSCHEME_ID[cscheme] if isinstance(cscheme, str) else cscheme
        prm.grad_method = GRAD_ID[grad_method] if isinstance(grad_method, str) else grad_method
        prm.limiter = LIMITER_ID[limiter] if isinstance(limiter, str) else limiter
        prm.tscheme = TSCHEME[tscheme] if isinstance(tscheme, str) else tscheme
        prm.timestep, prm.prtr, prm.viscos, prm.densit, prm.lowre = timestep, prtr, viscos, densit, int(lowre)
        rep = Report()
        lo, hi = C.c_double(0.0), C.c_double(0.0)
        check(lib().fcp_calcsc(self.h, C.byref(prm), field_id(phi), C.byref(rep), C.byref(lo), C.byref(hi)), "fcp_calcsc")
        return rep, lo.value, hi.value

    def grad_gauss_fvx(self, phi, grad):
        """Grad of the tensor-field layer: the two-pass Gauss gradient of fvxGradient.f90:1549-1662."""
        check(lib().fcp_grad_gauss_fvx(self.h, field_id(phi), field_id(grad)), "fcp_grad_gauss_fvx")

    def grad_gauss_iter(self, phi, grad, nigrad: int):
        """grad_gauss of the MPI tree (src-par/gradients.f90:1547-1664): `nigrad` passes of gradco."""
        check(lib().fcp_grad_gauss_iter(self.h, field_id(phi), field_id(grad), int(nigrad)), "fcp_grad_gauss_iter")

    def modify_viscosity_sgs(self, model, urfVis: float, viscos: float):
        """modify_viscosity_wale_sgs / modify_viscosity_vreman_sgs."""
        mid = {"wale": 0, "vreman": 1}[model] if isinstance(model, str) else int(model)
        check(lib().fcp_modify_viscosity_sgs(self.h, mid, urfVis, viscos), "fcp_modify_viscosity_sgs")

    def wall_distance(self):
        """wall_distance.f90:75-133 -> field WALLDIST; returns the ICCG report."""
        rep = Report()
        check(lib().fcp_wall_distance(self.h, C.byref(rep)), "fcp_wall_distance")
        return rep

    def calc_strain_and_vorticity(self):
        check(lib().fcp_calc_strain_and_vorticity(self.h), "fcp_calc_strain_and_vorticity")

    def modify_mu_eff_k_epsilon_rlzb(self, urfVis: float, viscos: float):
        check(lib().fcp_modify_mu_eff_k_epsilon_rlzb(self.h, urfVis, viscos), "fcp_modify_mu_eff_k_epsilon_rlzb")

    def modify_mu_eff_k_omega_sst(self, urfVis: float, viscos: float, densit: float, lowre: bool = False):
        check(lib().fcp_modify_mu_eff_k_omega_sst(self.h, urfVis, viscos, densit, int(lowre)), "fcp_modify_mu_eff_k_omega_sst")

    def update_boundary(self, field):
        """updateBoundary(phi), boundary/updateBoundary.f90."""
        check(lib().fcp_update_boundary(self.h, field_id(field)), "fcp_update_boundary")

    def calcuvw(self, solver="bicgstab", maxiter=5, tol_abs=1e-13, tol_rel=0.025, urf=(0.8, 0.8, 0.8), gds=1.0, cscheme="cds", grad_method="gauss",
                limiter="none", pscheme="linear", tscheme="steady", timestep=0.0, piso=False, const_mflux=False, gradPcmf=0.0, viscos=0.0):
        """calcuvw, Velocity/velocity.f90:50-750.  Returns the three solver reports (U, V, W)."""
        prm = UvwParams()
        prm.solver = SOLVER_ID[solver] if isinstance(solver, str) else solver
        prm.maxiter, prm.tol_abs, prm.tol_rel = maxiter, tol_abs, tol_rel
        prm.urf[0], prm.urf[1], prm.urf[2] = urf
        prm.gds = gds
        prm.cscheme = CSCHEME_ID[cscheme] if isinstance(cscheme, str) else cscheme
        prm.grad_method = GRAD_ID[grad_method] if isinstance(grad_method, str) else grad_method
        prm.limiter = LIMITER_ID[limiter] if isinstance(limiter, str) else limiter
        prm.pscheme = PSCHEME[pscheme] if isinstance(pscheme, str) else pscheme
        prm.tscheme = TSCHEME[tscheme] if isinstance(tscheme, str) else tscheme
        prm.piso, prm.timestep, prm.const_mflux, prm.gradPcmf, prm.viscos = int(piso), timestep, int(const_mflux), gradPcmf, viscos
        reps = (Report * 3)()
        check(lib().fcp_calcuvw(self.h, C.byref(prm), reps), "fcp_calcuvw")
        return [reps[i] for i in range(3)]

    def laplacian(self, mu, phi):
        check(lib().fcp_laplacian(self.h, field_id(mu), field_id(phi)), "fcp_laplacian")

    def gradp_and_sources(self, pscheme, p):
        ps = PSCHEME[pscheme] if isinstance(pscheme, str) else int(pscheme)
        check(lib().fcp_gradp_and_sources(self.h, ps, field_id(p)), "fcp_gradp_and_sources")

    def set_flux_variant(self, variant: int, grad_method: int = 0):
        """0: facefluxmass2 of the serial tree (default); 1: facefluxmass of the MPI tree on inner faces (quirk Q10), velocity gradients by `grad_method`."""
        check(lib().fcp_set_flux_variant(self.h, int(variant), int(grad_method)), "fcp_set_flux_variant")

    def assemble_pcorr_simple(self, const_mflux: bool = False, flomas: float = 0.0):
        check(lib().fcp_assemble_pcorr_simple(self.h, int(const_mflux), flomas), "fcp_assemble_pcorr_simple")

    def correct_simple(self, pscheme, urfp: float, pRefCell: int):
        ps = PSCHEME[pscheme] if isinstance(pscheme, str) else int(pscheme)
        check(lib().fcp_correct_simple(self.h, ps, urfp, pRefCell), "fcp_correct_simple")

    def nonorth_corrector(self):
        check(lib().fcp_nonorth_corrector(self.h), "fcp_nonorth_corrector")

    def calcp_simple(self, solver="iccg", maxiter=20, tol_abs=1e-13, tol_rel=0.025, urfp=0.3, npcor=1, pRefCell=1,
                     pscheme="linear", const_mflux=False, flomas=0.0, zero_pp=False):
        prm = SimpleParams(SOLVER_ID[solver] if isinstance(solver, str) else solver, maxiter, tol_abs, tol_rel, urfp, npcor, pRefCell,
                           PSCHEME[pscheme] if isinstance(pscheme, str) else pscheme, int(const_mflux), flomas, int(zero_pp))
        reps = (Report * max(npcor, 1))()
        check(lib().fcp_calcp_simple(self.h, C.byref(prm), reps), "fcp_calcp_simple")
        return [reps[i] for i in range(npcor)]

    # ---- multi-GPU -------------------------------------------------------------------------------------------
    def comm_init(self, rank: int, nranks: int, unique_id: bytes, peer_rank: np.ndarray):
        pr = np.ascontiguousarray(peer_rank, dtype=np.int32)
        buf = C.create_string_buffer(unique_id, 128)
        check(lib().fcp_comm_init(self.h, rank, nranks, buf, _i(pr)), "fcp_comm_init")

    def set_process_facint(self, fpro: np.ndarray):
        """src-par/geometry.f90:822-868: the interpolation factors of the process faces (patch order), e.g. the line-plane variant of the MPI tree
        (`mesh.facint_line_plane` on the global mesh, carried into `Mesh.fpro` by `mesh.partition`); overrides what fcp_comm_init computed."""
        f = np.ascontiguousarray(fpro, dtype=np.float64)
        check(lib().fcp_set_process_facint(self.h, _d(f), int(f.size)), "fcp_set_process_facint")

    def set_process_orientation(self, flipped: np.ndarray):
        """per process face (patch order): non-zero when this rank's cell is the face's NEIGHBOUR in the unpartitioned mesh (`mesh.process_face_flipped`);
        needed by the k-omega SST pair on a partition, which takes 1/sigma from the face's owner cell (k_omega_SST.f90:440-449)."""
        f = np.ascontiguousarray(flipped, dtype=np.int32)
        check(lib().fcp_set_process_orientation(self.h, _i(f), int(f.size)), "fcp_set_process_orientation")

    def comm_mode(self) -> str:
        """'p2p' (peer-memory stores over NVLink fused into the kernels), 'nccl' (send/recv + all-gather) or 'none'."""
        return {1: "p2p", 0: "nccl"}.get(int(lib().fcp_comm_mode(self.h)), "none")

    def exchange(self, field):
        check(lib().fcp_exchange(self.h, field_id(field)), "fcp_exchange")

    def global_sum(self, v: float) -> float:
        x = C.c_double(v)
        check(lib().fcp_global_sum(self.h, C.byref(x)), "fcp_global_sum")
        return x.value

    def global_max(self, v: float) -> float:
        x = C.c_double(v)
        check(lib().fcp_global_max(self.h, C.byref(x)), "fcp_global_max")
        return x.value

    def global_min(self, v: float) -> float:
        x = C.c_double(v)
        check(lib().fcp_global_min(self.h, C.byref(x)), "fcp_global_min")
        return x.value

    def global_isum(self, i: int) -> int:
        x = C.c_int64(int(i))
        check(lib().fcp_global_isum(self.h, C.byref(x)), "fcp_global_isum")
        return int(x.value)

    # ---- timing ----------------------------------------------------------------------------------------------
    def profile_enable(self, on: bool = True):
        check(lib().fcp_profile_enable(self.h, int(on)))

    def profile_reset(self):
        check(lib().fcp_profile_reset(self.h))

    def profile_read(self):
        """kernel class -> (total device ms, launches) since the last reset (CUDA events around every launch)."""
        out = {}
        for i, name in enumerate(KERNEL_CLASSES):
            ms, cnt = C.c_double(), C.c_int64()
            check(lib().fcp_profile_read(self.h, i, C.byref(ms), C.byref(cnt)))
            if cnt.value:
                out[name] = (ms.value, int(cnt.value))
        return out

    def timer_start(self):
        check(lib().fcp_timer_start(self.h))

    def timer_stop(self) -> float:
        ms = C.c_float()
        check(lib().fcp_timer_stop(self.h, C.byref(ms)))
        return float(ms.value)

    def flush_l2(self):
        check(lib().fcp_flush_l2(self.h))


def comm_unique_id() -> bytes:
    buf = C.create_string_buffer(128)
    check(lib().fcp_comm_unique_id(buf), "fcp_comm_unique_id")
    return buf.raw


def report_line(rep: Report, chvar: str) -> str:
    buf = C.create_string_buffer(256)
    check(lib().fcp_report_line(C.byref(rep), chvar.encode(), buf, 256))
    return buf.value.decode()


class CsrSolver:
    """Explicit-CSR signature of the reference's dpcg/iccg/bicgstab(n,nnz,ia,ja,a,diag,fi,rhs,...)."""

    def __init__(self, ia, ja, diag, device: int = 0):
        self.ia = np.ascontiguousarray(ia, np.int32)
        self.ja = np.ascontiguousarray(ja, np.int32)
        self.diag = np.ascontiguousarray(diag, np.int32)
        self.n = self.ia.size - 1
        self.h = C.c_void_p()
        check(lib().fcp_solver_create(self.n, self.ja.size, _i(self.ia), _i(self.ja), _i(self.diag), device, C.byref(self.h)), "fcp_solver_create")

    def solve(self, solver, a, fi, rhs, itr_max, tol_abs, tol_rel) -> Report:
        a = np.ascontiguousarray(a, np.float64)
        rhs = np.ascontiguousarray(rhs, np.float64)
        assert fi.dtype == np.float64 and fi.flags.c_contiguous
        rep = Report()
        sid = SOLVER_ID[solver] if isinstance(solver, str) else int(solver)
        check(lib().fcp_solver_solve(self.h, sid, _d(a), _d(fi), _d(rhs), itr_max, tol_abs, tol_rel, C.byref(rep)), "fcp_solver_solve")
        return rep

    def close(self):
        if getattr(self, "h", None) is not None and self.h:
            lib().fcp_solver_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

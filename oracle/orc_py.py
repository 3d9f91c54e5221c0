"""ctypes binding of the CPU oracle (oracle/liborc.so).  TEST INFRASTRUCTURE ONLY: imported by
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs, never by the
product package.  Accepts any mesh object exposing the attribute names of the reference's
``geometry`` module (see freecappuccino-dev_b200/mesh.py:Mesh)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import List

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

SUM_SEQ, SUM_TREE = 0, 1
DPCG, ICCG, BICGSTAB, GAUSS_SEIDEL = 1, 2, 3, 4

_pd = C.POINTER(C.c_double)
_pi = C.POINTER(C.c_int32)


class OrcMesh(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("numCells", "numInnerFaces", "numBoundaryFaces", "numFaces", "numTotal", "numBoundaries")] + \
               [("owner", _pi), ("neighbour", _pi)] + \
               [(n, _pd) for n in ("arx", "ary", "arz", "xf", "yf", "zf", "facint", "Df", "xc", "yc", "zc", "vol")] + \
               [(n, _pi) for n in ("bctype", "nfaces", "startFace", "iBndValueStart", "startFaceTwin")]


class OrcReport(C.Structure):
    _fields_ = [("res0", C.c_double), ("resl", C.c_double), ("factor", C.c_double), ("resor", C.c_double), ("iters", C.c_int32)]


class OrcRank(C.Structure):
    _fields_ = [("mesh", C.POINTER(OrcMesh)), ("ia", _pi), ("ja", _pi), ("diag", _pi), ("a", _pd), ("apr", _pd),
                ("npro", C.c_int32), ("peer_rank", _pi), ("peer_patch", _pi), ("fi", _pd), ("rhs", _pd)]


def build(force: bool = False, name: str = "liborc.so") -> str:
    so = os.path.join(_HERE, name)
    src = [os.path.join(_HERE, f) for f in ("orc.cpp", "orc.h")]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in src if os.path.exists(s)):
        subprocess.check_call(["make", "-C", _HERE, "-B", name], stdout=subprocess.DEVNULL)
    return so


def use_openmp(threads: int = 0) -> int:
    """Switch this process to liborc_omp.so (the multi-threaded timing build; bench.py only).  Returns the thread count."""
    global _LIB
    os.environ.setdefault("OMP_PROC_BIND", "close")
    if threads:
        os.environ["OMP_NUM_THREADS"] = str(threads)
    elif os.environ.get("OMP_NUM_THREADS") in (None, "", "1"):   # torchrun exports OMP_NUM_THREADS=1
        os.environ["OMP_NUM_THREADS"] = str(os.cpu_count() or 1)
    try:
        path = build(name="liborc_omp.so")
    except (subprocess.CalledProcessError, OSError):
        return 1                                  # no OpenMP toolchain here: stay on the serial library
    _LIB = None
    _load(path)
    return int(os.environ["OMP_NUM_THREADS"])


def _load(path: str):
    global _LIB
    _LIB = C.CDLL(path)
    _LIB.orc_small.restype = C.c_double
    _LIB.orc_sum_tree.restype = C.c_double
    _LIB.orc_sum_tree.argtypes = [_pd, C.c_int64]
    _LIB.orc_csr_nnz.restype = C.c_int32
    return _LIB


def lib():
    if _LIB is None:
        _load(build())
    return _LIB


def _d(a):
    assert a.dtype == np.float64 and a.flags.c_contiguous, (a.dtype, a.flags)
    return a.ctypes.data_as(_pd)


def _i(a):
    assert a.dtype == np.int32 and a.flags.c_contiguous
    return a.ctypes.data_as(_pi)


class MeshView:
    """Keeps the numpy arrays alive next to the C struct."""

    def __init__(self, m):
        self.m = m
        self.keep = dict(owner=np.ascontiguousarray(m.owner, np.int32), neighbour=np.ascontiguousarray(m.neighbour, np.int32),
                         bctype=np.ascontiguousarray(m.bctype, np.int32), nfaces=np.ascontiguousarray(m.nfaces, np.int32),
                         startFace=np.ascontiguousarray(m.startFace, np.int32),
                         iBndValueStart=np.ascontiguousarray(m.iBndValueStart, np.int32),
                         startFaceTwin=np.ascontiguousarray(m.twin_start() if hasattr(m, "twin_start") else np.full(m.numBoundaries, -1), np.int32))
        for n in ("arx", "ary", "arz", "xf", "yf", "zf", "facint", "Df", "xc", "yc", "zc", "vol"):
            self.keep[n] = np.ascontiguousarray(getattr(m, n), np.float64)
        s = OrcMesh()
        s.numCells, s.numInnerFaces, s.numBoundaryFaces = m.numCells, m.numInnerFaces, m.numBoundaryFaces
        s.numFaces, s.numTotal, s.numBoundaries = m.numFaces, m.numTotal, m.numBoundaries
        for n in ("owner", "neighbour", "bctype", "nfaces", "startFace", "iBndValueStart", "startFaceTwin"):
            setattr(s, n, _i(self.keep[n]))
        for n in ("arx", "ary", "arz", "xf", "yf", "zf", "facint", "Df", "xc", "yc", "zc", "vol"):
            setattr(s, n, _d(self.keep[n]))
        self.s = s

    @property
    def ptr(self):
        return C.byref(self.s)


def small() -> float:
    return lib().orc_small()


def sum_tree(v: np.ndarray) -> float:
    v = np.ascontiguousarray(v, np.float64)
    return lib().orc_sum_tree(_d(v), v.size)


def geometry(points, face_nodes, face_nnodes, owner, neighbour, numCells):
    nF, F = owner.shape[0], neighbour.shape[0]
    x, y, z = (np.ascontiguousarray(points[:, k]) for k in range(3))
    node = np.ascontiguousarray(face_nodes, np.int32)   # [nF, nomax] row-major == node(nomax, nF) column-major
    out = {n: np.zeros(nF) for n in ("arx", "ary", "arz", "xf", "yf", "zf")}
    out.update({n: np.zeros(numCells) for n in ("vol", "xc", "yc", "zc")})
    out.update({n: np.zeros(F) for n in ("facint", "Df")})
    fnn = np.ascontiguousarray(face_nnodes, np.int32)
    own = np.ascontiguousarray(owner, np.int32)
    nb = np.ascontiguousarray(neighbour, np.int32)
    lib().orc_geometry(C.c_int32(points.shape[0]), C.c_int32(numCells), C.c_int32(F), C.c_int32(nF), _d(x), _d(y), _d(z),
                       _i(fnn), _i(node), C.c_int32(node.shape[1]), _i(own), _i(nb),
                       *[_d(out[n]) for n in ("arx", "ary", "arz", "xf", "yf", "zf", "vol", "xc", "yc", "zc", "facint", "Df")])
    return out


class Csr:
    def __init__(self, mesh):
        mv = MeshView(mesh)
        self.mv = mv
        self.n = mesh.numCells
        self.nnz = lib().orc_csr_nnz(mv.ptr)
        self.ia = np.zeros(self.n + 1, np.int32)
        self.ja = np.zeros(self.nnz, np.int32)
        self.diag = np.zeros(self.n, np.int32)
        nper = lib().orc_num_periodic(mv.ptr)
        self.icell_jcell = np.zeros(mesh.numInnerFaces + nper, np.int32)      # sparse_matrix.f90:246-247
        self.jcell_icell = np.zeros(mesh.numInnerFaces + nper, np.int32)
        lib().orc_csr_create(mv.ptr, _i(self.ia), _i(self.ja), _i(self.diag), _i(self.icell_jcell), _i(self.jcell_icell))


def laplacian(mesh, csr: Csr, mu, phi, su):
    a = np.zeros(csr.nnz)
    lib().orc_laplacian(csr.mv.ptr, _i(csr.diag), _i(csr.icell_jcell), _i(csr.jcell_icell), _d(mu), _d(phi), _d(a), C.c_int32(csr.nnz), _d(su))
    return a


def grad_gauss(mesh, u):
    mv = MeshView(mesh)
    g = np.zeros((mesh.numTotal, 3))
    lib().orc_grad_gauss(mv.ptr, _d(u), _d(g))
    return g


def create_matrix_lsq(mesh, weighted: bool):
    mv = MeshView(mesh)
    D = np.zeros((mesh.numCells, 9))
    lib().orc_create_matrix_lsq(mv.ptr, C.c_int(int(weighted)), _d(D))
    return D


def grad_lsq(mesh, weighted: bool, Dmat, phi, row2_correct: bool = False):
    mv = MeshView(mesh)
    g = np.zeros((mesh.numTotal, 3))
    lib().orc_grad_lsq(mv.ptr, C.c_int(int(weighted)), C.c_int(int(row2_correct)), _d(Dmat), _d(phi), _d(g))
    return g


def gradp_and_sources(mesh, pscheme: int, p, apu, dPdxi):
    """p (numTotal) and dPdxi (numTotal,3) are updated in place; returns su, sv, sw."""
    mv = MeshView(mesh)
    su, sv, sw = (np.zeros(mesh.numCells) for _ in range(3))
    lib().orc_gradp_and_sources(mv.ptr, C.c_int(pscheme), _d(p), _d(apu), _d(su), _d(sv), _d(sw), _d(dPdxi))
    return su, sv, sw


def assemble_pcorr(mesh, csr: Csr, den, u, v, w, p, pp, dPdxi, apu, const_mflux=False, flomas=0.0, apv=None, apw=None):
    a = np.zeros(csr.nnz)
    su = np.zeros(mesh.numCells)
    flmass = np.zeros(mesh.numFaces)
    lib().orc_assemble_pcorr(csr.mv.ptr, _i(csr.diag), _i(csr.icell_jcell), _i(csr.jcell_icell), C.c_int32(csr.nnz),
                             _d(den), _d(u), _d(v), _d(w), _d(p), _d(pp), _d(dPdxi), _d(apu),
                             _d(apv) if apv is not None else None, _d(apw) if apw is not None else None,
                             C.c_int(int(const_mflux)), C.c_double(flomas), _d(a), _d(su), _d(flmass))
    return a, su, flmass


def assemble_pcorr_into(mesh, csr: Csr, den, u, v, w, p, pp, dPdxi, apu, a, su, flmass, const_mflux=False, flomas=0.0, apv=None, apw=None):
    lib().orc_assemble_pcorr(csr.mv.ptr, _i(csr.diag), _i(csr.icell_jcell), _i(csr.jcell_icell), C.c_int32(csr.nnz),
                             _d(den), _d(u), _d(v), _d(w), _d(p), _d(pp), _d(dPdxi), _d(apu),
                             _d(apv) if apv is not None else None, _d(apw) if apw is not None else None,
                             C.c_int(int(const_mflux)), C.c_double(flomas), _d(a), _d(su), _d(flmass))


def assemble_pcorr_mpi_into(mesh, csr: Csr, den, u, v, w, p, pp, dPdxi, apu, apv, apw, gU, gV, gW, a, su, flmass, const_mflux=False, flomas=0.0):
    """p' assembly with the MPI tree's inner-face flux (quirk Q10, src-par/faceflux_mass.f90:28-180); gU, gV, gW: (numTotal, 3) velocity gradients."""
    lib().orc_assemble_pcorr_mpi(csr.mv.ptr, _i(csr.diag), _i(csr.icell_jcell), _i(csr.jcell_icell), C.c_int32(csr.nnz),
                                 _d(den), _d(u), _d(v), _d(w), _d(p), _d(pp), _d(dPdxi), _d(apu), _d(apv), _d(apw),
                                 C.c_int(int(const_mflux)), C.c_double(flomas), _d(a), _d(su), _d(flmass), _d(gU), _d(gV), _d(gW))


def correct_simple(mesh, csr: Csr, pscheme, a, den, u, v, w, p, pp, apu, apv, apw, urfp, pRefCell, dPdxi, flmass):
    su, sv, sw = (np.zeros(mesh.numCells) for _ in range(3))
    lib().orc_correct_simple(csr.mv.ptr, _i(csr.icell_jcell), C.c_int(pscheme), _d(a), _d(den), _d(u), _d(v), _d(w), _d(p), _d(pp),
                             _d(apu), _d(apv), _d(apw), C.c_double(urfp), C.c_int32(pRefCell), _d(su), _d(sv), _d(sw), _d(dPdxi), _d(flmass))
    return su, sv, sw


def constant_mass_flow_forcing(mesh, magUbar, apu, u, sum_mode=0):
    """constant_mass_flow_forcing.f90: corrects u in place; returns (gragPplus, magUbarStar)."""
    mv = MeshView(mesh)
    f = lib().orc_constant_mass_flow_forcing
    f.restype = C.c_double
    ustar = C.c_double(0.0)
    g = f(mv.ptr, C.c_double(magUbar), _d(apu), _d(u), C.c_int(sum_mode), C.byref(ustar))
    return g, ustar.value


def update_boundary(mesh, phi):
    """boundary/updateBoundary.f90, in place."""
    mv = MeshView(mesh)
    lib().orc_update_boundary(mv.ptr, _d(phi))
    return phi


def nonorth_corrector(mesh, den, apu, dPdxi, su, flmass):
    mv = MeshView(mesh)
    lib().orc_nonorth_corrector(mv.ptr, _d(den), _d(apu), _d(dPdxi), _d(su), _d(flmass))


def spmv(ia, ja, a, x):
    y = np.zeros(ia.size - 1)
    lib().orc_spmv(C.c_int32(ia.size - 1), _i(ia), _i(ja), _d(a), _d(x), _d(y))
    return y


def solve(solver: int, ia, ja, a, diag, fi, rhs, itr_max, tol_abs, tol_rel, sum_mode=SUM_SEQ) -> OrcReport:
    """fi is updated in place (first n entries)."""
    n, nnz = ia.size - 1, ja.size
    rep = OrcReport()
    fn = {DPCG: lib().orc_dpcg, ICCG: lib().orc_iccg, BICGSTAB: lib().orc_bicgstab, GAUSS_SEIDEL: lib().orc_gauss_seidel}[solver]
    fn(C.c_int32(n), C.c_int32(nnz), _i(ia), _i(ja), _d(a), _i(diag), _d(fi), _d(rhs), C.c_int32(itr_max),
       C.c_double(tol_abs), C.c_double(tol_rel), C.c_int(sum_mode), C.byref(rep))
    return rep


def report_line(solver: int, chvar: str, rep: OrcReport) -> str:
    buf = C.create_string_buffer(256)
    lib().orc_report_line(C.c_int(solver), chvar.encode(), C.byref(rep), buf, C.c_int(256))
    return buf.value.decode()


def dpcg_par(parts: List, csrs: List[Csr], a_list, apr_list, fi_list, rhs_list, itr_max, tol_abs, tol_rel, sum_mode=SUM_SEQ) -> OrcReport:
    P = len(parts)
    ranks = (OrcRank * P)()
    keep = []
    for r in range(P):
        mv = csrs[r].mv
        pr = np.ascontiguousarray(parts[r].peer_rank, np.int32)
        pp = np.ascontiguousarray(parts[r].peer_patch, np.int32)
        apr = np.ascontiguousarray(apr_list[r], np.float64) if apr_list[r].size else np.zeros(1)
        keep += [pr, pp, apr]
        ranks[r].mesh = C.pointer(mv.s)
        ranks[r].ia, ranks[r].ja, ranks[r].diag = _i(csrs[r].ia), _i(csrs[r].ja), _i(csrs[r].diag)
        ranks[r].a, ranks[r].apr, ranks[r].npro = _d(a_list[r]), _d(apr), parts[r].npro
        ranks[r].peer_rank, ranks[r].peer_patch = _i(pr), _i(pp)
        ranks[r].fi, ranks[r].rhs = _d(fi_list[r]), _d(rhs_list[r])
    rep = OrcReport()
    lib().orc_dpcg_par(C.c_int32(P), ranks, C.c_int32(itr_max), C.c_double(tol_abs), C.c_double(tol_rel), C.c_int(sum_mode), C.byref(rep))
    return rep


LIM_NONE, LIM_BJ, LIM_VENKAT, LIM_R3, LIM_MDL = 0, 1, 2, 3, 4


def slope_limiter(mesh, csr: Csr, kind: int, phi, dPhidxi):
    """gradients.f90:288-656; dPhidxi (numTotal,3) is limited in place."""
    lib().orc_slope_limiter(csr.mv.ptr, _i(csr.ia), _i(csr.ja), _i(csr.diag), C.c_int(kind), _d(phi), _d(dPhidxi))
    return dPhidxi


def create_matrix_lsq_qr(mesh):
    mv = MeshView(mesh)
    D = np.zeros((mesh.numCells, 6, 3))          # D(3,6,numCells) column-major
    lib().orc_create_matrix_lsq_qr.restype = C.c_int
    rc = lib().orc_create_matrix_lsq_qr(mv.ptr, _d(D))
    if rc != 0:
        raise ValueError("create_matrix_lsq_qr: a cell has more than 6 faces (gradients.f90:924 m=6)")
    return D


def grad_lsq_qr(mesh, D, phi):
    mv = MeshView(mesh)
    g = np.zeros((mesh.numTotal, 3))
    lib().orc_grad_lsq_qr.restype = C.c_int
    rc = lib().orc_grad_lsq_qr(mv.ptr, _d(D), _d(phi), _d(g))
    if rc != 0:
        raise ValueError("grad_lsq_qr: a cell has more than 6 faces")
    return g


def calcp_piso(mesh, csr: Csr, solver, maxiter, tol_abs, tol_rel, sum_mode, ncorr, npcor, pscheme, urfp, const_mflux, flomas,
               rU, rV, rW, den, apu, apv, apw, a, u, v, w, p, pp, dPdxi, flmass):
    """a (momentum coefficients in, pressure matrix out), u, v, w, p, pp, dPdxi, flmass are updated in place.
    Returns (reports, su, sv, sw, h)."""
    n = mesh.numCells
    su, sv, sw = (np.zeros(n) for _ in range(3))
    h = np.zeros(csr.nnz)
    reps = (OrcReport * (ncorr * npcor))()
    lib().orc_calcp_piso(csr.mv.ptr, _i(csr.ia), _i(csr.ja), _i(csr.diag), _i(csr.icell_jcell), _i(csr.jcell_icell), C.c_int32(csr.nnz),
                         C.c_int(solver), C.c_int32(maxiter), C.c_double(tol_abs), C.c_double(tol_rel), C.c_int(sum_mode),
                         C.c_int(ncorr), C.c_int(npcor), C.c_int(pscheme), C.c_double(urfp), C.c_int(int(const_mflux)), C.c_double(flomas),
                         _d(rU), _d(rV), _d(rW), _d(den), _d(apu), _d(apv), _d(apw), _d(a), _d(h), _d(u), _d(v), _d(w), _d(p), _d(pp),
                         _d(su), _d(sv), _d(sw), _d(dPdxi), _d(flmass), reps)
    return [reps[i] for i in range(ncorr * npcor)], su, sv, sw, h


CSCHEMES = ["cds", "central", "linearUpwind", "kappa", "muscl", "umist", "koren", "smart", "avl-smart", "charm", "vanleer", "ospre", "minmod",
            "boundedLinearUpwind", "boundedLinearUpwind02", "boundedCentral", "fromm", "cui", "quick", "spl13"]


class OrcUvwParams(C.Structure):
    _fields_ = [("solver", C.c_int32), ("maxiter", C.c_int32), ("tol_abs", C.c_double), ("tol_rel", C.c_double), ("urf", C.c_double * 3),
                ("gds", C.c_double), ("cscheme", C.c_int32), ("grad_method", C.c_int32), ("limiter", C.c_int32), ("pscheme", C.c_int32),
                ("tscheme", C.c_int32), ("timestep", C.c_double), ("piso", C.c_int32), ("const_mflux", C.c_int32), ("gradPcmf", C.c_double),
                ("viscos", C.c_double), ("sum_mode", C.c_int32), ("pad", C.c_int32)]


def face_value(mesh, cscheme: int, ijp: int, ijn: int, xf, yf, zf, lam, u, dUdxi) -> float:
    mv = MeshView(mesh)
    lib().orc_face_value.restype = C.c_double
    return lib().orc_face_value(mv.ptr, C.c_int(cscheme), C.c_int32(ijp), C.c_int32(ijn), C.c_double(xf), C.c_double(yf), C.c_double(zf),
                                C.c_double(lam), _d(u), _d(dUdxi))


def calcuvw(mesh, csr: Csr, prm: OrcUvwParams, f: dict, a: np.ndarray):
    """The momentum predictor (velocity.f90:50-750).  f: u,v,w,p (numTotal, updated in place), den, vis (numTotal), visw
    (numBoundaryFaces), flmass (numFaces), apu (numTotal, in: the weighted pscheme reads it), optionally uo..wooo.  a: stale
    matrix values in, the W-equation matrix out.  Returns a dict of everything the routine leaves behind."""
    n, nT = mesh.numCells, mesh.numTotal
    out = {k: np.zeros(n) for k in ("su", "sv", "sw", "spu", "spv", "sp", "rU", "rV", "rW")}
    out.update({k: np.zeros(nT) for k in ("apv", "apw")})
    out["apu"] = f["apu"]
    out.update({k: np.zeros((nT, 3)) for k in ("dUdxi", "dVdxi", "dWdxi", "dPdxi")})
    reps = (OrcReport * 3)()
    opt = lambda k: _d(f[k]) if k in f and f[k] is not None else None  # noqa: E731
    lib().orc_calcuvw(csr.mv.ptr, _i(csr.ia), _i(csr.ja), _i(csr.diag), _i(csr.icell_jcell), _i(csr.jcell_icell), C.c_int32(csr.nnz), C.byref(prm),
                      _d(f["u"]), _d(f["v"]), _d(f["w"]), _d(f["p"]), _d(f["den"]), _d(f["vis"]), _d(f["visw"]), _d(f["flmass"]),
                      opt("uo"), opt("vo"), opt("wo"), opt("uoo"), opt("voo"), opt("woo"), opt("uooo"), opt("vooo"), opt("wooo"),
                      _d(a), _d(out["su"]), _d(out["sv"]), _d(out["sw"]), _d(out["spu"]), _d(out["spv"]), _d(out["sp"]),
                      _d(out["apu"]), _d(out["apv"]), _d(out["apw"]), _d(out["dUdxi"]), _d(out["dVdxi"]), _d(out["dWdxi"]), _d(out["dPdxi"]),
                      _d(out["rU"]), _d(out["rV"]), _d(out["rW"]), reps)
    out["reps"] = [reps[i] for i in range(3)]
    return out


# ---- row f4: scalar transport -----------------------------------------------------------------------------------------------
SC_GENERIC, SC_TKE_RLZB, SC_EPS_RLZB, SC_TKE_SST, SC_OMEGA_SST = 0, 1, 2, 3, 4


class OrcScalarParams(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("kind", "solver", "maxiter", "cscheme", "grad_method", "limiter", "tscheme", "sum_mode")] + \
               [(n, C.c_double) for n in ("tol_abs", "tol_rel", "urf", "gds", "timestep", "prtr", "viscos", "densit")]


def calcsc(mesh, csr: Csr, prm: OrcScalarParams, f: dict):
    """The calcsc template (k_epsilon_rlzb.f90:52-790 + scalar_fluxes.f90).  f: te, ed (numTotal; the one prm.kind selects is solved
    for and updated in place; kind 0 solves f['phi']), den, vis (numTotal), visw, dnw (numBoundaryFaces), flmass (numFaces), u, v, w,
    magStrain (numCells), optionally phio/phioo, su_vol/sp_vol (kind 0).  Returns a dict with a, su, sp, grad, gen, tau, rep, fimin, fimax."""
    n, nT, B = mesh.numCells, mesh.numTotal, mesh.numBoundaryFaces
    phi = f["phi"] if prm.kind == SC_GENERIC else (f["te"] if prm.kind in (SC_TKE_RLZB, SC_TKE_SST) else f["ed"])
    out = dict(a=np.zeros(csr.nnz), su=np.zeros(n), sp=np.zeros(n), grad=np.zeros((nT, 3)), tau=np.zeros(max(B, 1)))
    out["gen"] = f["gen"] if prm.kind == SC_OMEGA_SST else np.zeros(n)        # the omega equation reads the production the k call left behind
    rep = OrcReport()
    fmin, fmax = C.c_double(0.0), C.c_double(0.0)
    opt = lambda k: _d(f[k]) if k in f and f[k] is not None else None  # noqa: E731
    lib().orc_calcsc(csr.mv.ptr, _i(csr.ia), _i(csr.ja), _i(csr.diag), _i(csr.icell_jcell), _i(csr.jcell_icell), C.c_int32(csr.nnz), C.byref(prm),
                     _d(phi), opt("phio"), opt("phioo"), opt("te"), opt("ed"), _d(f["den"]), _d(f["vis"]), opt("visw"), opt("dnw"), _d(f["flmass"]),
                     opt("u"), opt("v"), opt("w"), opt("magStrain"), _d(out["gen"]), _d(out["tau"]), opt("su_vol"), opt("sp_vol"),
                     _d(out["a"]), _d(out["su"]), _d(out["sp"]), _d(out["grad"]), C.byref(rep), C.byref(fmin), C.byref(fmax),
                     opt("fsst"), opt("walldist"), opt("dTEdxi"), C.c_int(int(f.get("lowre", 0))))
    out.update(rep=rep, fimin=fmin.value, fimax=fmax.value)
    return out


def wall_geometry(mesh):
    """geometry.f90:698-754: dnw, srdw per wall face and dns, srds per symmetry face, in patch order."""
    mv = MeshView(mesh)
    nw = max(int(mesh.nfaces[np.asarray(mesh.bctype) == 0].sum()), 1)
    ns = max(int(mesh.nfaces[np.asarray(mesh.bctype) == 3].sum()), 1)
    dnw, srdw, dns, srds = np.zeros(nw), np.zeros(nw), np.zeros(ns), np.zeros(ns)
    lib().orc_wall_geometry(mv.ptr, _d(dnw), _d(srdw), _d(dns), _d(srds))
    return dnw, srdw, dns, srds


def modify_mu_eff_sst(mesh, urf, viscos, densit, lowre, magStrain, walldist, te, ed, den, u, v, w, dnw, vis, visw):
    """modify_mu_eff of k_omega_SST.f90:790-958: vis and visw updated in place; returns ypl, tau."""
    mv = MeshView(mesh)
    B = max(mesh.numBoundaryFaces, 1)
    ypl, tau = np.zeros(B), np.zeros(B)
    lib().orc_modify_mu_eff_sst(mv.ptr, C.c_double(urf), C.c_double(viscos), C.c_double(densit), C.c_int(int(lowre)), _d(magStrain), _d(walldist),
                                _d(te), _d(ed), _d(den), _d(u), _d(v), _d(w), _d(dnw), _d(vis), _d(visw), _d(ypl), _d(tau))
    return ypl, tau


def calc_strain_and_vorticity(mesh, dUdxi, dVdxi, dWdxi):
    mv = MeshView(mesh)
    s, w = np.zeros(mesh.numCells), np.zeros(mesh.numCells)
    lib().orc_calc_strain_and_vorticity(mv.ptr, _d(dUdxi), _d(dVdxi), _d(dWdxi), _d(s), _d(w))
    return s, w


def modify_mu_eff_rlzb(mesh, urf, viscos, dUdxi, dVdxi, dWdxi, te, ed, den, u, v, w, dnw, vis, visw):
    """modify_mu_eff (k_epsilon_rlzb.f90:792-975): vis (numTotal) and visw (numBoundaryFaces) updated in place; returns ypl, tau."""
    mv = MeshView(mesh)
    B = max(mesh.numBoundaryFaces, 1)
    ypl, tau = np.zeros(B), np.zeros(B)
    lib().orc_modify_mu_eff_rlzb(mv.ptr, C.c_double(urf), C.c_double(viscos), _d(dUdxi), _d(dVdxi), _d(dWdxi), _d(te), _d(ed), _d(den),
                                 _d(u), _d(v), _d(w), _d(dnw), _d(vis), _d(visw), _d(ypl), _d(tau))
    return ypl, tau


def grad_gauss_fvx(mesh, u):
    """fvxGradient.f90:1549-1662: the two-pass Gauss gradient with the gradco skewness correction (SoA result, numCells)."""
    mv = MeshView(mesh)
    gx, gy, gz = (np.zeros(mesh.numCells) for _ in range(3))
    lib().orc_grad_gauss_fvx(mv.ptr, _d(u), _d(gx), _d(gy), _d(gz))
    return gx, gy, gz


def grad_gauss_iter(mesh, u, nigrad):
    """src-par/gradients.f90:1547-1664: the MPI tree's grad_gauss, `nigrad` passes of gradco (SoA result, numCells)."""
    mv = MeshView(mesh)
    gx, gy, gz = (np.zeros(mesh.numCells) for _ in range(3))
    lib().orc_grad_gauss_iter(mv.ptr, _d(u), C.c_int(int(nigrad)), _d(gx), _d(gy), _d(gz))
    return gx, gy, gz


SGS_WALE, SGS_VREMAN = 0, 1


def modify_viscosity_sgs(mesh, model, urf, viscos, u, v, w, den, vis, visw):
    """wale_sgs.f90 / vremanSGS.f90: vis (numTotal) and visw (numBoundaryFaces) updated in place."""
    mv = MeshView(mesh)
    lib().orc_modify_viscosity_sgs(mv.ptr, C.c_int(model), C.c_double(urf), C.c_double(viscos), _d(u), _d(v), _d(w), _d(den), _d(vis), _d(visw))

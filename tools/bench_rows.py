#!/usr/bin/env python
"""Per-kernel device timings of the rows beyond the headline step (f1 momentum predictor, a9 PISO, f3 periodic, f4 scalar transport and
turbulence closures, the gradient family and the limiters) on one B200, against their algorithmic bytes.

    python tools/bench_rows.py [--n 256] [--reps 5] [--periodic]

One JSON line per operation: device ms of the whole C-ABI call (CUDA events on the library stream, fcp_timer_*), the per-kernel-class times
the library's profiler attributes inside it (fcp_profile_*: an event pair around every launch of the class), the algorithmic bytes of the
dominant kernel and the resulting GB/s and fraction of the measured HBM peak.  bench.py stays the contract benchmark (the SIMPLE pressure
step); this script exists so that ONE gpurun call yields a roofline figure for every other kernel family.  Fields are synthetic and smooth;
solver iterations are capped at 2 because only the assembly kernels are of interest here."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402  (synthetic_fields, hbm_peak)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", "--cells", dest="n", type=int, default=256)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--periodic", action="store_true", help="periodic x and z (channel395-style) instead of the closed cavity")
    ap.add_argument("--only", default="", help="comma-separated substring filter on the operation names (A/B runs of a few kernels)")
    args = ap.parse_args()
    import fcb200  # noqa: F401
    from fcb200 import lib as L
    from fcb200 import mesh as M
    n = args.n
    xs = np.linspace(0.0, 1.0, n + 1)
    t0 = time.perf_counter()
    m = M.hex_mesh_fast(xs, xs, xs, dict(left="empty", right="periodic", back="empty", front="periodic") if args.periodic else None)
    f = bench.synthetic_fields(m)
    N, F, B, nT = m.numCells, m.numInnerFaces, m.numBoundaryFaces, m.numTotal
    ctx = L.Context(m, 0)
    nnz = ctx.nnz
    pi = np.pi
    f["vis"] = 0.01 * m.boundary_values_of(lambda x, y, z: 1.0 + 0.5 * np.sin(pi * x) * np.sin(pi * y))
    f["te"] = 0.02 * m.boundary_values_of(lambda x, y, z: 1.0 + 0.5 * np.sin(2 * x) * np.cos(3 * y))
    f["ed"] = 0.5 * m.boundary_values_of(lambda x, y, z: 1.0 + 0.4 * np.cos(x + 2 * y))
    for k in ("u", "v", "w", "p", "den", "apu", "apv", "apw", "vis", "te", "ed"):
        ctx.upload(k.upper(), f[k])
    visw = np.zeros(nT); visw[N:] = 0.01
    dnw = np.zeros(nT); dnw[N:] = 0.5 / n
    wd = np.zeros(nT); wd[:N] = 0.5 / n + np.minimum(m.yc[:N], 1.0 - m.yc[:N])
    ctx.upload("VISW", visw); ctx.upload("DNW", dnw); ctx.upload("WALLDIST", wd)
    own, nb = m.owner[:F].astype(np.int64) - 1, m.neighbour.astype(np.int64) - 1
    flm = np.zeros(m.numFaces)
    flm[:F] = 0.5 * ((f["u"][own] + f["u"][nb]) * m.arx[:F] + (f["v"][own] + f["v"][nb]) * m.ary[:F] + (f["w"][own] + f["w"][nb]) * m.arz[:F])
    ctx.upload("FLMASS", flm)
    ctx.upload("A", np.zeros(nnz))
    for k in "UVW":
        ctx.copy(k + "O", k); ctx.copy(k + "OO", k)
    ctx.copy("PHIO", "TE"); ctx.copy("PHIOO", "TE")
    ctx.sync()
    setup_s = time.perf_counter() - t0
    peak, peak_src = bench.hbm_peak()
    sc = dict(solver="bicgstab", maxiter=2, tol_abs=1e-30, tol_rel=1e-30, urf=0.7, gds=1.0, cscheme="muscl", limiter="Venkatakrishnan", viscos=1e-3, densit=1.0)

    ops = [
        # name, callable, profiler class of the dominant kernel, its algorithmic bytes per launch, launches of that class per call
        ("grad_gauss", lambda: ctx.grad(L.GRAD_GAUSS, "P", "G0"), "grad", 40 * F + 40 * N + 36 * B),
        ("grad_lsq", lambda: ctx.grad(L.GRAD_LSQ, "P", "G0"), "grad", 8 * F + 128 * N + 36 * B),
        ("grad_gauss_fvx (2 passes)", lambda: ctx.grad_gauss_fvx("P", "G0"), "grad", 2 * (80 * F + 88 * N + 36 * B)),
        ("limiter Venkatakrishnan", lambda: ctx.slope_limiter("Venkatakrishnan", "P", "G0"), "limiter", 4 * nnz + 68 * N),
        ("calcuvw (assembly kernel)", lambda: ctx.calcuvw(solver="bicgstab", maxiter=2, tol_abs=1e-30, tol_rel=1e-30, cscheme="muscl", limiter="Venkatakrishnan",
                                                          tscheme="bdf2", timestep=0.01, piso=True, viscos=1e-3), "uvw", 136 * F + 400 * N),
        ("calcp_piso (H(U) kernel)", lambda: ctx.calcp_piso(solver="dpcg", maxiter=2, tol_abs=1e-30, tol_rel=1e-30, ncorr=1, npcor=1), "piso_h", 12 * nnz + 100 * N),
        ("calcsc k (realizable)", lambda: ctx.calcsc("TE", kind="tke_rlzb", prtr=1.0, **sc), "scalar", 112 * F + 144 * N + 36 * B),
        ("calcsc epsilon (realizable)", lambda: ctx.calcsc("ED", kind="eps_rlzb", prtr=1 / 1.2, **sc), "scalar", 112 * F + 144 * N + 36 * B),
        ("calcsc k (SST)", lambda: ctx.calcsc("TE", kind="tke_sst", **sc), "scalar", 112 * F + 152 * N + 36 * B),
        ("calcsc omega (SST)", lambda: ctx.calcsc("ED", kind="omega_sst", **sc), "scalar", 112 * F + 184 * N + 36 * B),
        ("modify_mu_eff (realizable)", lambda: ctx.modify_mu_eff_k_epsilon_rlzb(0.8, 1e-3), None, 112 * N),
        ("modify_mu_eff (SST)", lambda: ctx.modify_mu_eff_k_omega_sst(0.8, 1e-3, 1.0), None, 56 * N),
        ("modify_viscosity_sgs WALE (6 gradient passes + algebra)", lambda: ctx.modify_viscosity_sgs("wale", 0.8, 1e-3), None, 6 * (80 * F + 88 * N + 36 * B) + 96 * N),
        ("modify_viscosity_sgs Vreman", lambda: ctx.modify_viscosity_sgs("vreman", 0.8, 1e-3), None, 6 * (80 * F + 88 * N + 36 * B) + 96 * N),
        ("constant_mass_flow_forcing", lambda: ctx.constant_mass_flow_forcing(0.1335, 0.0), None, 40 * N),
        ("update_boundary", lambda: ctx.update_boundary("TE"), None, 24 * B),
    ]
    ctx.create_lsq_grad_matrix(L.GRAD_LSQ)
    ctx.calc_strain_and_vorticity()
    for name, fn, klass, nbytes in ops:
        if args.only and not any(t and t in name for t in args.only.split(",")):
            continue
        fn()                                        # warm-up (allocates lazily created fields, builds level schedules ...)
        ctx.sync()
        ctx.profile_enable(True); ctx.profile_reset()
        ms = []
        for _ in range(args.reps):
            ctx.timer_start()
            fn()
            ms.append(ctx.timer_stop())
        prof = ctx.profile_read()
        ctx.profile_enable(False)
        line = dict(op=name, n=n, periodic=bool(args.periodic), call_ms=float(np.median(ms)), classes={k: dict(total_ms=v[0], launches=v[1]) for k, v in prof.items()},
                    algorithmic_bytes=int(nbytes), peak_gbs=peak, peak_source=peak_src)
        if klass and klass in prof:
            tot, cnt = prof[klass]
            per_call = tot / args.reps
            line.update(kernel_class=klass, kernel_ms_per_call=per_call, achieved_gbs=nbytes / (per_call * 1e-3) / 1e9, frac=nbytes / (per_call * 1e-3) / 1e9 / peak)
        else:
            line.update(achieved_gbs=nbytes / (line["call_ms"] * 1e-3) / 1e9, frac=nbytes / (line["call_ms"] * 1e-3) / 1e9 / peak,
                        note="whole call (several small kernels + host synchronisation where the call returns a scalar)")
        print(json.dumps(line), flush=True)
    print(json.dumps(dict(setup_s=setup_s, cells=N, faces=F, nnz=nnz)), flush=True)
    ctx.close()


if __name__ == "__main__":
    main()

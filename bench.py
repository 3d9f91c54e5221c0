#!/usr/bin/env python
"""bench.py -- one SIMPLE pressure step of the freeCappuccino hot path on N B200s of one node.

Workload (BASELINE.json configs[2]): synthetic 3-D lid-driven-cavity mesh, n^3 uniform hexahedra (default 256^3 =
16.7 M cells), split in z-slabs over the ranks (src-par layout).  One "step" = what calcp_simple does per outer
iteration on this path:
    gradp_and_sources(p)                       (Pressure/nablap.f90)            gradient kernel
    p'-equation assembly, facefluxmass2        (calcp_simple.f90:69-118)        assembly kernel
    csrsolve('dpcg', pp, su, tolRel=1e-8)      (linear_solvers.f90:206-359)     SpMV+dot / axpy kernels
    flux, velocity and pressure correction     (calcp_simple.f90:331-429)
Inputs are reset to the same synthetic state before every step (otherwise the second step would start converged).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--n 256] [--impl ours|reference]
Multi-GPU: python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 ... bench.py --gpus N
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FALLBACK_HBM_GBS = 6650.0   # /opt/skills/guides/B200_PROFILING.md fallback when MEASURED_PEAKS.json is absent
INPUT_FIELDS = ("u", "v", "w", "p", "den", "apu", "apv", "apw")
OUTPUT_FIELDS = ("u", "v", "w", "p", "pp")
TOL_REL = 1e-8
MAXITER = 20000


def synthetic_fields(m):
    """u = sin(pi x) cos(pi y), v = -cos(pi x) sin(pi y), w = 0, p = cos(2 pi x) cos(2 pi y)/4,
    apu = apv = apw = 0.8/(6 nu h), den = 1.  (u,v) is not discretely divergence free, so the p' system has a real RHS."""
    pi = np.pi
    h = float(m.vol[: m.numCells].mean()) ** (1.0 / 3.0)
    f = {}
    f["u"] = m.boundary_values_of(lambda x, y, z: np.sin(pi * x) * np.cos(pi * y) * (1.0 + 0.5 * np.sin(pi * z)))
    f["v"] = m.boundary_values_of(lambda x, y, z: -np.cos(pi * x) * np.sin(pi * y))
    f["w"] = m.boundary_values_of(lambda x, y, z: 0.1 * np.sin(pi * x) * np.sin(2 * pi * z))
    f["p"] = m.boundary_values_of(lambda x, y, z: 0.25 * np.cos(2 * pi * x) * np.cos(2 * pi * y))
    f["den"] = np.ones(m.numTotal)
    ap = 0.8 / (6.0 * 0.01 * h)
    for k in ("apu", "apv", "apw"):
        f[k] = np.full(m.numTotal, ap)
    return f


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region (B200_PROFILING.md)."""

    def __init__(self, gpu_index=0):
        vis = [v for v in os.environ.get("CUDA_VISIBLE_DEVICES", "").split(",") if v.strip()]
        if vis and gpu_index < len(vis) and vis[gpu_index].strip().isdigit():
            gpu_index = int(vis[gpu_index])          # nvidia-smi numbers the physical GPUs
        self.tmp = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=self.tmp, stderr=subprocess.DEVNULL)
        except OSError:
            self.proc = None

    def stop(self):
        out = dict(sm_mhz=None, sm_max_mhz=None, reasons=[], samples=0)
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        self.tmp.flush()
        self.tmp.seek(0)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.tmp.read().splitlines():
            t = [x.strip() for x in line.split(",")]
            if len(t) < 7:
                continue
            try:
                sm.append(float(t[0])); mx.append(float(t[1]))
            except ValueError:
                continue
            for nm, v in zip(names, t[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        self.tmp.close()
        os.unlink(self.tmp.name)
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), reasons=sorted(reasons), samples=len(sm))
        return out


def hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            with open(p) as fh:
                d = json.load(fh)
            if "hbm_gbs" in d:
                return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except (OSError, ValueError):
            pass
    return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


def l2_note(working_set_bytes):
    """timing hygiene note of `config.l2` (both arms print the same string for the same mesh)"""
    if working_set_bytes > 2 * 126e6:
        return f"inputs larger than L2 (matrix + vectors = {working_set_bytes / 1e6:.0f} MB per rank vs 126 MB L2)"
    return "working set fits L2: L2-resident numbers"


def pinned_copy(a):
    """numpy view of a pinned torch buffer holding a copy of `a` (H2D copies from it are true DMA transfers)."""
    import torch
    t = torch.empty(a.shape, dtype=torch.float64, pin_memory=True)
    v = t.numpy()
    v[...] = a
    return t, v


# ------------------------------------------------------------------------------------------------------------------
def oracle_step_time(m, f, gpu_iters, max_sample_iters, threads=1):
    """Times the CPU restatement (oracle/, -O2 -ffp-contract=off) on the SAME mesh and inputs: gradp + assembly +
    correction in full, DPCG for `max_sample_iters` iterations, and extrapolates the solve to `gpu_iters`, the iteration
    count of the converged run (0: the sample's own count).  threads = 1: the serial reference stand-in; threads > 1:
    liborc_omp.so, the DPCG loops and the SpMV split over the host cores like the reference's src-par MPI build (the face
    loops stay serial: they are < 1 % of the step)."""
    from oracle import orc_py as O
    if threads > 1:
        threads = O.use_openmp(threads)
    t0 = time.perf_counter()
    c = O.Csr(m)
    t_csr = time.perf_counter() - t0
    g = {k: v.copy() for k, v in f.items()}
    g["pp"] = np.zeros(m.numTotal)
    dP = np.zeros((m.numTotal, 3))
    t0 = time.perf_counter()
    O.gradp_and_sources(m, 0, g["p"], g["apu"], dP)
    t_gradp = time.perf_counter() - t0
    a = np.zeros(c.nnz); su = np.zeros(m.numCells); flm = np.zeros(m.numFaces)
    t0 = time.perf_counter()
    O.assemble_pcorr_into(m, c, g["den"], g["u"], g["v"], g["w"], g["p"], g["pp"], dP, g["apu"], a, su, flm)
    t_asm = time.perf_counter() - t0
    if 0 < max_sample_iters < MAXITER:
        # a bounded sample: the first call touches the solver's work arrays for the first time (page faults, cold caches) and over-stated the
        # converged solve's time per iteration by 11-20 % in round 1 -- run the sample twice from the same start and time the second run
        pp0 = g["pp"].copy()
        O.solve(O.DPCG, c.ia, c.ja, a, c.diag, g["pp"], su, max_sample_iters, 1e-30, TOL_REL)
        g["pp"][:] = pp0
    t0 = time.perf_counter()
    rep = O.solve(O.DPCG, c.ia, c.ja, a, c.diag, g["pp"], su, max_sample_iters, 1e-30, TOL_REL)
    t_solve = time.perf_counter() - t0
    its = max(rep.iters, 1)
    converged = rep.resl / (rep.res0 + 1e-300) < TOL_REL
    t0 = time.perf_counter()
    O.correct_simple(m, c, 0, a, g["den"], g["u"], g["v"], g["w"], g["p"], g["pp"], g["apu"], g["apv"], g["apw"], 0.3, 1, dP, flm)
    t_corr = time.perf_counter() - t0
    t_iter = t_solve / (its + 0.5)       # the initial residual costs about half an iteration
    n_it = its if (converged or not gpu_iters) else gpu_iters
    total = t_gradp + t_asm + t_corr + t_iter * (n_it + 0.5)
    return dict(ms=1e3 * total, t_gradp=t_gradp, t_asm=t_asm, t_corr=t_corr, t_iter=t_iter, sample_iters=its, iters_used=n_it,
                converged=bool(converged), t_csr=t_csr, measured_s=t_gradp + t_asm + t_solve + t_corr, threads=threads)


def run_reference(args, rank, world):
    """--impl reference: the reference's own CPU implementation of the path on the box's host cores.  The Fortran
    reference cannot be built in this image (no gfortran/flang/nvfortran; probed), so this is the C++ restatement
    (oracle/) = kind 'port', run with all host threads (OpenMP over the DPCG loops and the SpMV = the src-par MPI build's
    work split).  The FIRST warm-up step runs the solve to convergence (that fixes the iteration count the workload
    needs: 1006 at 256^3, the same as the GPU arm); every later step is a bounded sample: the face loops in full and
    --ref-iters DPCG iterations, extrapolated to that count."""
    if rank != 0:
        return
    from fcb200 import mesh as M
    n = args.n
    m = M.hex_mesh_fast(*(np.linspace(0.0, 1.0, n + 1),) * 3)
    f = synthetic_fields(m)
    threads = args.ref_threads or (os.cpu_count() or 1)
    full = oracle_step_time(m, f, 0, MAXITER, threads)              # converged solve: fixes the iteration count AND the time per iteration
    count = full["iters_used"]
    t_iter_full = full["t_iter"]
    times = []
    for s in range(max(args.warmup - 1, 0) + args.steps):
        r = oracle_step_time(m, f, count, min(args.ref_iters, count), threads)
        # the step's value: its own face loops (measured in full) + the solve at the per-iteration time of the CONVERGED solve (a 40-iteration
        # sample runs colder than the 1006-iteration solve and over-estimated the step by 11-20 % in round 1)
        r["ms"] = 1e3 * (r["t_gradp"] + r["t_asm"] + r["t_corr"] + t_iter_full * (count + 0.5))
        if s >= max(args.warmup - 1, 0):
            times.append(r)
    ms = float(np.mean([t["ms"] for t in times]))
    r = times[-1]
    faces_s = r["t_gradp"] + r["t_asm"] + r["t_corr"]
    sample = (f"{n}^3 mesh, C++ restatement of the Fortran path on {r['threads']} host threads (OpenMP over the DPCG loops and the SpMV = the src-par work split; "
              f"the face loops are serial: {faces_s:.1f} s = {100 * faces_s / (ms / 1e3):.0f} % of the step): the first warm-up step is the WHOLE step with the solve run to "
              f"convergence ({count} DPCG iterations, {full['ms'] / 1e3:.1f} s, {t_iter_full * 1e3:.1f} ms/iteration); each timed step = gradp_and_sources + assembly + "
              f"correction in full ({r['t_gradp']:.2f}+{r['t_asm']:.2f}+{r['t_corr']:.2f} s) + a {r['sample_iters']}-iteration DPCG sample ({r['t_iter'] * 1e3:.1f} ms/iteration, "
              f"cross-check only); value = face loops of the step + {count} iterations at the converged solve's time per iteration "
              f"(one fully measured step: full_step_ms); {r['measured_s']:.1f} s of CPU work per timed step")
    line = dict(metric=f"SIMPLE iter time ({n}^3 hex cavity: gradp + p' assembly + DPCG to 1e-8 + correction)", value=ms, unit="ms",
                impl="reference", n_gpus=args.gpus, steps=args.steps, warmup=args.warmup, ms_per_step=ms, higher_is_better=False,
                scaling="strong", vs_baseline=None, dtype="f64", data="synthetic",
                config=dict(workload=f"synthetic 3D lid-driven cavity {n}^3 hex ({n**3} cells), pressure PCG", solver="dpcg", tol_rel=TOL_REL,
                            pcg_iters=count, partition=f"z-slabs x{world}", comm="none" if world == 1 else "host threads",
                            l2=l2_note(12 * (m.numCells + 2 * m.numInnerFaces) + 60 * m.numCells)),
                cpu_baseline=dict(value=ms, unit="ms", cores=r["threads"], kind="port", sample=sample,
                                  full_step_ms=full["ms"], ms_per_iteration_converged=1e3 * t_iter_full, ms_per_iteration_sample=1e3 * r["t_iter"]),
                e2e=dict(value=ms, unit="ms", h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------------------------
def run_poly(args, rank, local_rank, world, torch, dist, L, M):
    """--workload poly (BASELINE.json configs[4]): the brick-pattern polyhedral mesh (10-faced cells, ~20 M at --poly-n 342) in z-slabs over the
    ranks; one step = what the reference's wall-distance / gradient pipeline does on such a mesh (mesh/wall_distance.f90:96-133 with the
    gradient methods of examples/elbow3D/input-1.nml):
        grad_gauss(phi)                          gradients.f90:1607-1693      k_grad_gauss      40 F + 40 N + 36 B bytes
        grad_lsq(phi)                            gradients.f90:782-893        k_grad_lsq         8 F + 128 N + 36 B bytes
        laplacian(mu, phi) with su = vol         fvImplicit/laplacian.f90     k_laplacian
        csrsolve('iccg', tolRel 1e-8)            linear_solvers.f90:364-545   IC(0) factor, level-scheduled sweeps, SpMV+dot, updates
    Across ranks the IC(0) preconditioner is block-Jacobi like src-par/iccg.f90, so the iteration count grows with the rank count."""
    nx = args.poly_n
    t_setup = time.perf_counter()
    m = M.polyhedral_partition_fast(nx, world, rank) if world > 1 else M.polyhedral_mesh_fast(nx)
    N0, F0, B0 = m.numCells, m.numInnerFaces, m.numBoundaryFaces
    phi = m.boundary_values_of(lambda x, y, z: np.sin(2.0 * x) * np.cos(3.0 * y) + z * z)
    vol_su = np.concatenate([m.vol[:N0], np.zeros(B0)])
    pin_phi, pin_su = pinned_copy(phi), pinned_copy(vol_su)
    out_g1, out_g2, out_x = pinned_copy(np.zeros((m.numTotal, 3))), pinned_copy(np.zeros((m.numTotal, 3))), pinned_copy(np.zeros(m.numTotal))
    ctx = L.Context(m, local_rank)
    if world > 1:
        uid = [L.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        ctx.comm_init(rank, world, uid[0], m.peer_rank)
    ctx.upload("S0", pin_phi[1])
    ctx.fill("VIS", -1.0)
    ctx.fill("S1", 0.0)
    ctx.create_lsq_grad_matrix(L.GRAD_LSQ)
    nnz0 = ctx.nnz + ctx.npro
    MAXIT = 5000

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()

    def step():
        ctx.grad(L.GRAD_GAUSS, "S0", "G0")
        ctx.grad(L.GRAD_LSQ, "S0", "G1")
        ctx.copy("SU", "S2")                      # su = vol (the laplacian call subtracts the boundary terms from it)
        ctx.laplacian("VIS", "S1")
        ctx.fill("PP", 0.0)
        return ctx.csrsolve(args.solver if args.solver != "dpcg" or args.poly_dpcg else "iccg", "PP", "SU", MAXIT, 1e-30, TOL_REL)

    def e2e_step():
        ctx.upload("S0", pin_phi[1])
        ctx.upload("S2", pin_su[1])
        rep = step()
        for fld, buf in (("G0", out_g1), ("G1", out_g2)):
            L.check(L.lib().fcp_field_download(ctx.h, L.field_id(fld), L._d(buf[1]), 3 * m.numTotal))
        L.check(L.lib().fcp_field_download(ctx.h, L.field_id("PP"), L._d(out_x[1]), m.numTotal))
        return rep
    ctx.upload("S2", pin_su[1])
    rep = None
    for _ in range(max(args.warmup, 1)):
        rep = step()
    ctx.sync()
    if rep.iters >= MAXIT or not np.isfinite(rep.resl):
        raise SystemExit(f"bench.py: the ICCG solve did not converge ({rep.iters} iterations, final residual {rep.resl})")
    t_setup = time.perf_counter() - t_setup
    ctx.profile_enable(False)
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    launches0 = L.launch_count()
    dev_ms = []
    for _ in range(args.steps):
        ctx.timer_start()
        rep = step()
        dev_ms.append(ctx.timer_stop())
    barrier()
    launches = L.launch_count() - launches0
    # profiled pass: each gradient kernel on its own, then the solve
    ctx.profile_enable(True)
    prof = {}
    for name, fn in (("grad_gauss", lambda: ctx.grad(L.GRAD_GAUSS, "S0", "G0")), ("grad_lsq", lambda: ctx.grad(L.GRAD_LSQ, "S0", "G1"))):
        ctx.profile_reset()
        for _ in range(5):
            fn()
        prof[name] = ctx.profile_read().get("grad")
    ctx.profile_reset()
    ctx.copy("SU", "S2"); ctx.laplacian("VIS", "S1"); ctx.fill("PP", 0.0)
    ctx.csrsolve("iccg", "PP", "SU", MAXIT, 1e-30, TOL_REL)
    prof.update(ctx.profile_read())
    ctx.profile_enable(False)
    e2e_step()
    barrier()
    e2e0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    barrier()
    e2e_ms = 1e3 * (time.perf_counter() - e2e0) / args.steps
    clocks = sampler.stop() if sampler else None
    stats = dict(ms=float(np.mean(dev_ms)), e2e=e2e_ms, cells=N0, launches=launches)
    allstats = [stats]
    if dist is not None:
        allstats = [None] * world
        dist.all_gather_object(allstats, stats)
    if rank != 0:
        ctx.close()
        if dist is not None:
            dist.destroy_process_group()
        return
    ms, e2e = max(q["ms"] for q in allstats), max(q["e2e"] for q in allstats)
    peak, peak_src = hbm_peak()

    def kernel_line(name, nbytes):
        if not prof.get(name):
            return None
        tot, cnt = prof[name]
        gbs = nbytes / (tot / cnt * 1e-3) / 1e9
        return dict(kernel=name, launches=cnt, avg_ms=tot / cnt, bytes_per_launch=nbytes, achieved_gbs=gbs, frac=gbs / peak)
    kl = {"grad_gauss": kernel_line("grad_gauss", 40 * F0 + 40 * N0 + 36 * B0), "grad_lsq": kernel_line("grad_lsq", 8 * F0 + 128 * N0 + 36 * B0),
          "laplacian": kernel_line("laplacian", 64 * F0 + 60 * N0), "precond": kernel_line("precond", 24 * F0 + 80 * N0),
          "spmv_dot": kernel_line("spmv_dot", 12 * nnz0 + 20 * N0), "cg_pk": kernel_line("cg_pk", 24 * N0), "cg_update": kernel_line("cg_update", 48 * N0),
          "dot": kernel_line("dot", 16 * N0)}
    gg = kl["grad_gauss"]
    total_cells = int(sum(q["cells"] for q in allstats))
    line = dict(
        metric=f"polyhedral step time (~{total_cells / 1e6:.1f} M ten-faced polyhedra: grad_gauss + grad_lsq + laplacian + ICCG to 1e-8)", value=ms, unit="ms",
        n_gpus=world, steps=args.steps, warmup=args.warmup, ms_per_step=ms, higher_is_better=False, scaling="strong", vs_baseline=None,
        dtype="f64", data="synthetic",
        config=dict(workload=f"synthetic polyhedral elbow3D-style mesh {nx}^3/2 ({total_cells} cells), Gauss/LSQ gradients + ICCG", solver="iccg",
                    tol_rel=TOL_REL, iccg_iters=int(rep.iters), partition=f"z-slabs x{world}", comm=ctx.comm_mode(),
                    l2=l2_note(12 * nnz0 + 60 * N0)),
        e2e=dict(value=e2e, unit="ms", h2d_bytes_per_step=16 * m.numTotal, d2h_bytes_per_step=56 * m.numTotal),
        gpu_launches=int(sum(q["launches"] for q in allstats)),
        roofline=dict(bound="hbm", achieved=gg["achieved_gbs"], peak=peak, unit="GB/s", frac=gg["frac"], traffic=None, kernel="k_grad_gauss (40 F + 40 N + 36 B)",
                      peak_source=peak_src, avg_launch_ms=gg["avg_ms"], launches=gg["launches"]) if gg else None,
        cpu_baseline=None, clocks=clocks, kernels={k: v for k, v in kl.items() if v},
        iccg=dict(iters=int(rep.iters), res0=rep.res0, resl=rep.resl), setup_s=t_setup)
    print(json.dumps(line), flush=True)
    ctx.close()
    if dist is not None:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--n", "--cells", dest="n", type=int, default=256, help="cells per direction of the cavity mesh (--cells under torchrun, whose own parser rejects --n as ambiguous)")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--solver", default="dpcg", choices=["dpcg", "iccg", "bicgstab"])
    ap.add_argument("--workload", default="cavity", choices=["cavity", "poly"], help="cavity: BASELINE configs[2] (the contract benchmark); poly: configs[4]")
    ap.add_argument("--poly-n", type=int, default=342, help="--workload poly: hexahedra per direction before merging (342 -> 20.0 M polyhedra)")
    ap.add_argument("--poly-dpcg", action="store_true", help="--workload poly: honour --solver dpcg (default: ICCG, the solver config 5 names)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--prof-steps", type=int, default=2, help="steps of the separate per-kernel profiled pass (CUDA events around every launch)")
    ap.add_argument("--cpu-iters", type=int, default=40, help="DPCG iterations timed by the cpu_baseline leg")
    ap.add_argument("--ref-iters", type=int, default=40, help="DPCG iterations timed per step by --impl reference (bounded sample)")
    ap.add_argument("--ref-threads", type=int, default=0, help="host threads of --impl reference / cpu_baseline (0: all cores)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import fcb200  # noqa: F401
    from fcb200 import lib as L
    from fcb200 import mesh as M
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group(backend="gloo", init_method="env://")
    if world != args.gpus and rank == 0:
        print(f"bench.py: WORLD_SIZE={world} but --gpus {args.gpus}; using {world}", file=sys.stderr)
    n = args.n
    if args.workload == "poly":
        return run_poly(args, rank, local_rank, world, torch, dist, L, M)

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()

    # ---- mesh partition of this rank, context, communicator ---------------------------------------------------------
    t_setup = time.perf_counter()
    m = M.block_partition_mesh((n, n, n), M.block_dims(world), rank)
    f = synthetic_fields(m)
    pinned = {k: pinned_copy(f[k]) for k in INPUT_FIELDS}
    out_pinned = {k: pinned_copy(np.zeros(m.numTotal)) for k in OUTPUT_FIELDS}
    ctx = None

    def make_context():
        """context + communicator + device-resident inputs; called again when the communication path has to be changed"""
        nonlocal ctx
        ctx = L.Context(m, local_rank)
        if world > 1:
            uid = [L.comm_unique_id() if rank == 0 else None]
            dist.broadcast_object_list(uid, src=0)
            ctx.comm_init(rank, world, uid[0], m.peer_rank)
        for k in INPUT_FIELDS:
            ctx.upload(k.upper(), pinned[k][1])
        # pristine copies of the fields a step overwrites
        for src, dst in (("U", "S0"), ("V", "S1"), ("W", "S2"), ("P", "S3")):
            ctx.copy(dst, src)
        ctx.sync()

    def reset_device_inputs():
        for src, dst in (("U", "S0"), ("V", "S1"), ("W", "S2"), ("P", "S3")):
            ctx.copy(src, dst)

    def step():
        ctx.gradp_and_sources("linear", "P")
        return ctx.calcp_simple(solver=args.solver, maxiter=MAXITER, tol_abs=1e-30, tol_rel=TOL_REL, urfp=0.3, npcor=1, pRefCell=1 if rank == 0 else 0,
                                zero_pp=True)[0]

    def e2e_step():
        for k in INPUT_FIELDS:
            ctx.upload(k.upper(), pinned[k][1])
        rep = step()
        probe = (0, m.numCells // 2, m.numCells - 1)
        for k in OUTPUT_FIELDS:                       # sentinels instead of clearing 0.7 GB of host memory inside the timed region:
            out_pinned[k][1][list(probe)] = np.nan    # a download that did not happen leaves them behind
        for k in OUTPUT_FIELDS:
            L.check(L.lib().fcp_field_download(ctx.h, L.field_id(k.upper()), L._d(out_pinned[k][1]), m.numTotal))
        for k in OUTPUT_FIELDS:
            if not np.isfinite(out_pinned[k][1][list(probe)]).all():
                raise SystemExit(f"bench.py: the e2e download of {k} did not overwrite the host buffer")
        return rep

    # ---- set-up + warm-up ----------------------------------------------------------------------------------------------
    # A multi-GPU run whose peer-memory path fails (a wait timed out) or does not converge (wrong ghosts) is repeated ONCE over the NCCL
    # path, all ranks together, and says so in `config.comm`: a slower valid number beats none (round 1 lost its 4- and 8-GPU runs this way).
    comm_note = ""
    for attempt in (0, 1):
        problem = None
        try:
            make_context()
            rep = None
            for _ in range(max(args.warmup, 1)):
                reset_device_inputs()
                rep = step()
            ctx.sync()
            if rep.iters >= MAXITER or not np.isfinite(rep.resl):
                problem = f"the solve did not converge ({rep.iters} iterations, final residual {rep.resl})"
        except L.FcpError as ex:
            problem = str(ex)
        problems = [problem]
        if dist is not None:
            problems = [None] * world
            dist.all_gather_object(problems, problem)
        bad = [f"rank {r}: {q}" for r, q in enumerate(problems) if q]
        if not bad:
            break
        mode = ctx.comm_mode() if ctx is not None else "none"
        if attempt == 1 or world == 1 or mode != "p2p" or os.environ.get("FCP_COMM") == "p2p":
            raise SystemExit("bench.py: " + "; ".join(bad))
        if rank == 0:
            print("bench.py: peer-memory path failed (" + "; ".join(bad) + "); repeating over NCCL", file=sys.stderr)
        comm_note = " (fallback: the peer-memory path failed in the warm-up)"
        try:
            ctx.close()
        except Exception:
            pass
        os.environ["FCP_COMM"] = "nccl"
    comm_mode = ctx.comm_mode() + comm_note
    t_setup = time.perf_counter() - t_setup

    # ---- timed: device-resident inputs (value).  The per-kernel event profiler is OFF here (round 1 timed `value` with two events around each of
    # ~3000 launches per step, about 6 ms per step); the kernel table comes from a separate profiled pass below. ------------------------------------
    ctx.profile_enable(False)
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    launches0 = L.launch_count()
    dev_ms = []
    wall0 = time.perf_counter()
    for _ in range(args.steps):
        reset_device_inputs()
        ctx.timer_start()
        rep = step()
        dev_ms.append(ctx.timer_stop())
    barrier()
    wall_ms = 1e3 * (time.perf_counter() - wall0)
    launches = L.launch_count() - launches0

    # ---- profiled pass (untimed for `value`): CUDA events around every launch of a kernel class, on the library's stream, same steps ------------------
    prof_steps = max(1, min(args.steps, args.prof_steps))
    ctx.profile_enable(True)
    ctx.profile_reset()
    barrier()
    prof_ms = []
    for _ in range(prof_steps):
        reset_device_inputs()
        ctx.timer_start()
        step()
        prof_ms.append(ctx.timer_stop())
    barrier()
    prof = ctx.profile_read()
    ctx.profile_enable(False)

    # ---- timed: end to end through the C-ABI with host buffers ----------------------------------------------------------------
    e2e_step()   # warm the transfer path
    barrier()
    e2e0 = time.perf_counter()
    for _ in range(args.steps):
        rep_e = e2e_step()
    barrier()
    e2e_ms = 1e3 * (time.perf_counter() - e2e0) / args.steps
    clocks = sampler.stop() if sampler else None

    # ---- reduce over ranks -----------------------------------------------------------------------------------------------------
    ms_local = float(np.mean(dev_ms))
    stats = dict(ms=ms_local, e2e=e2e_ms, cells=m.numCells, nnz=ctx.nnz, faces=m.numInnerFaces, launches=launches, prof=prof)
    if dist is not None:
        allstats = [None] * world
        dist.all_gather_object(allstats, stats)
    else:
        allstats = [stats]
    if rank != 0:
        ctx.close()
        if dist is not None:
            dist.destroy_process_group()
        return
    ms = max(s["ms"] for s in allstats)
    e2e = max(s["e2e"] for s in allstats)
    iters = int(rep.iters)
    peak, peak_src = hbm_peak()

    # roofline of the dominant kernel (SpMV + p.Ap dot): algorithmic bytes 12 nnz + 20 N per launch, rank 0's share
    def kernel_line(name, bytes_per_launch):
        if name not in prof:
            return None
        tot, cnt = prof[name]
        avg_ms = tot / cnt
        gbs = bytes_per_launch / (avg_ms * 1e-3) / 1e9
        return dict(kernel=name, launches=cnt, avg_ms=avg_ms, total_ms=tot, bytes_per_launch=bytes_per_launch, achieved_gbs=gbs, frac=gbs / peak)
    N0, nnz0, F0, B0 = m.numCells, ctx.nnz + ctx.npro, m.numInnerFaces, m.numBoundaryFaces
    kl = {
        "spmv_dot": kernel_line("spmv_dot", 12 * nnz0 + 20 * N0),
        "cg_pk": kernel_line("cg_pk", (32 if args.solver == "dpcg" else 24) * N0),            # dpcg: res, adiag, pk r + pk w; iccg: zk, pk r + pk w
        "cg_update": kernel_line("cg_update", (56 if args.solver == "dpcg" else 48) * N0),    # fi, res r+w, pk, zk (+ adiag for the Jacobi z)
        "assemble": kernel_line("assemble", 80 * F0 + 124 * N0),
        # two launches per step: gradp_and_sources(p) [40 F + 64 N + 36 B] and the same kernel on pp with the velocity / pressure correction of
        # calcp_simple.f90:416-429 fused in [+ 120 N: u, v, w r+w, apu, apv, apw, p r+w, pp -- SURVEY 8d row a6]; bytes = their mean
        "gradp": kernel_line("gradp", 40 * F0 + 64 * N0 + 36 * B0 + 60 * N0),
        "correct_flux": kernel_line("correct_flux", 36 * F0 + 8 * N0),
        "precond": kernel_line("precond", 24 * F0 + 80 * N0),       # IC(0)/ILU(0) apply: both triangles once + ia, diag, d, r, z
    }
    dom = kl["spmv_dot"]
    dom_name = "k_spmv_dot_pipe<1,false,8> (SELL-32 SpMV + p.Ap dot)"
    if args.solver != "dpcg" and kl["precond"] and kl["precond"]["total_ms"] > dom["total_ms"]:
        dom, dom_name = kl["precond"], "k_precond_apply (level-scheduled IC(0)/ILU(0) forward + backward sweep)"
    traffic = None    # dram bytes per launch of the dominant kernel from the committed ncu --set full capture (same mesh size, 1 GPU only)
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as fh:
            tj = json.load(fh)
        if world == 1 and tj.get("n") == n:
            traffic = tj["dram_bytes_per_launch"]
    except (OSError, ValueError, KeyError):
        pass
    roofline = dict(bound="hbm", achieved=dom["achieved_gbs"], peak=peak, unit="GB/s", frac=dom["frac"], traffic=traffic,
                    kernel=dom_name, peak_source=peak_src,
                    share_of_step=dom["total_ms"] / (float(np.mean(prof_ms)) * prof_steps), avg_launch_ms=dom["avg_ms"], launches=dom["launches"],
                    measured=f"CUDA events around every launch on the library stream, separate pass of {prof_steps} steps right after the timed region "
                             f"({float(np.mean(prof_ms)):.1f} ms/step with the events on vs {ms:.1f} ms/step without)")
    # DPCG iteration as a whole: 12 nnz + 108 N ideal bytes
    it_ms = sum(kl[k]["total_ms"] for k in ("spmv_dot", "cg_pk", "cg_update") if kl[k]) / max(dom["launches"], 1)
    it_gbs = (12 * nnz0 + 108 * N0) / (it_ms * 1e-3) / 1e9

    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline:
        threads = args.ref_threads or (os.cpu_count() or 1)
        r = oracle_step_time(m, f, iters, args.cpu_iters, threads)
        cpu_baseline = dict(value=r["ms"], unit="ms", cores=r["threads"], kind="port",
                            sample=(f"same {n}^3 mesh and inputs, C++ restatement (oracle/, the Fortran reference cannot be built here) on {r['threads']} host "
                                    f"threads (OpenMP over the DPCG loops and the SpMV, like the src-par MPI build; face loops serial): gradp + assembly + correction "
                                    f"in full ({r['t_gradp']:.2f}+{r['t_asm']:.2f}+{r['t_corr']:.2f} s), DPCG {r['sample_iters']} iterations timed after an untimed warm run of the same sample "
                                    f"({r['t_iter'] * 1e3:.1f} ms/iter), extrapolated to the {iters} iterations of the converged solve; {r['measured_s']:.1f} s measured"))
    h2d = 8 * m.numTotal * len(INPUT_FIELDS)
    d2h = 8 * m.numTotal * len(OUTPUT_FIELDS)
    line = dict(
        metric=f"SIMPLE iter time ({n}^3 hex cavity: gradp + p' assembly + DPCG to 1e-8 + correction)", value=ms, unit="ms",
        n_gpus=world, steps=args.steps, warmup=args.warmup, ms_per_step=ms, higher_is_better=False, scaling="strong", vs_baseline=None,
        dtype="f64", data="synthetic",
        config=dict(workload=f"synthetic 3D lid-driven cavity {n}^3 hex ({n**3} cells), pressure PCG", solver=args.solver, tol_rel=TOL_REL,
                    pcg_iters=iters, partition=f"z-slabs x{world}", comm=comm_mode, l2=l2_note(12 * nnz0 + 60 * N0)),
        e2e=dict(value=e2e, unit="ms", h2d_bytes_per_step=h2d, d2h_bytes_per_step=d2h),
        gpu_launches=int(sum(s["launches"] for s in allstats)),
        roofline=roofline, cpu_baseline=cpu_baseline, clocks=clocks,
        kernels={k: v for k, v in kl.items() if v},
        pcg=dict(iters=iters, res0=rep.res0, resl=rep.resl, iteration_ms=it_ms, iteration_gbs_ideal_bytes=it_gbs, iteration_frac=it_gbs / peak),
        wall_ms_per_step=wall_ms / args.steps, setup_s=t_setup,
    )
    print(json.dumps(line), flush=True)
    ctx.close()
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""A/B of the face-kernel switches on one B200, all variants inside ONE process and ONE context (GPU minutes are scarce).

    python tools/face_ab.py [--n 256] [--reps 5] [--out gpurun_out/r02_face_ab.txt]

The library reads FCP_FACE_OCC / FCP_FACE_PF / FCP_FACE_CL / FCP_FACE_OG / FCP_ASM_W at every launch of a face kernel, so a variant is selected by setting the
environment between calls.  Per variant: the inputs are restored, every operation runs once and a fingerprint of its results (wrap-around sum
of the 64-bit patterns) is compared with the first variant's -- a variant that changes a bit is flagged, not timed --, then `reps` timed calls
with the library's per-class profiler (CUDA events around the launches of the class).  Operations: grad_gauss, grad_lsq (class grad),
gradp_and_sources (class gradp: the plain launch), calcp_simple with 2 solver iterations (classes assemble and gradp: the launch with the fused
velocity / pressure correction).  Prints one table and the best variant per kernel as `export` lines (sourced by tools/gpu_session_face.sh)."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402  (synthetic_fields, hbm_peak; does not import torch at module level)


def fingerprint(a: np.ndarray) -> int:
    return int(np.ascontiguousarray(a).view(np.uint64).sum(dtype=np.uint64))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=256)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--out", default="")
    ap.add_argument("--round", type=int, default=3, help="which set of variants (see the code)")
    args = ap.parse_args()
    import fcb200  # noqa: F401
    from fcb200 import lib as L
    from fcb200 import mesh as M
    n = args.n
    t0 = time.perf_counter()
    xs = np.linspace(0.0, 1.0, n + 1)
    m = M.hex_mesh_fast(xs, xs, xs)
    f = bench.synthetic_fields(m)
    N, F, B = m.numCells, m.numInnerFaces, m.numBoundaryFaces
    ctx = L.Context(m, 0)
    for k in bench.INPUT_FIELDS:
        ctx.upload(k.upper(), f[k])
    for src, dst in (("U", "S0"), ("V", "S1"), ("W", "S2"), ("P", "S3")):
        ctx.copy(dst, src)
    ctx.create_lsq_grad_matrix(L.GRAD_LSQ)
    ctx.sync()
    setup_s = time.perf_counter() - t0
    peak, _ = bench.hbm_peak()
    nbytes = dict(fvx=2 * (80 * F + 88 * N + 36 * B), grad_gauss=40 * F + 40 * N + 36 * B, grad_lsq=8 * F + 128 * N + 36 * B, gradp_plain=40 * F + 64 * N + 36 * B,
                  gradp_fused=40 * F + 64 * N + 36 * B + 120 * N, assemble=80 * F + 124 * N)

    def restore():
        for src, dst in (("U", "S0"), ("V", "S1"), ("W", "S2"), ("P", "S3")):
            ctx.copy(src, dst)

    def simple():
        return ctx.calcp_simple(solver="dpcg", maxiter=2, tol_abs=1e-30, tol_rel=1e-30, urfp=0.3, npcor=1, pRefCell=1, zero_pp=True)

    def run_ops(record):
        """every operation once (record = fingerprints) or args.reps times (record = per-class ms)"""
        out = {}
        reps = 1 if record == "fp" else args.reps
        for name in ("grad_gauss", "grad_lsq", "fvx", "gradp_plain", "simple"):
            restore()
            ctx.sync()
            if record == "ms":
                ctx.profile_enable(True); ctx.profile_reset()
            for _ in range(reps):
                if name == "grad_gauss":
                    ctx.grad(L.GRAD_GAUSS, "P", "G0")
                elif name == "grad_lsq":
                    ctx.grad(L.GRAD_LSQ, "P", "G0")
                elif name == "fvx":
                    ctx.grad_gauss_fvx("P", "G0")
                elif name == "gradp_plain":
                    ctx.gradp_and_sources("linear", "P")
                else:
                    simple()
            if record == "ms":
                prof = ctx.profile_read()
                ctx.profile_enable(False)
                if name == "simple":
                    out["assemble"] = prof["assemble"][0] / max(prof["assemble"][1], 1)
                    out["gradp_fused"] = prof["gradp"][0] / max(prof["gradp"][1], 1)
                else:
                    klass = "gradp" if name == "gradp_plain" else "grad"
                    out[name] = prof[klass][0] / max(prof[klass][1], 1)
            else:
                if name in ("grad_gauss", "grad_lsq", "fvx"):
                    out[name] = fingerprint(ctx.download("G0"))
                elif name == "gradp_plain":
                    out[name] = fingerprint(ctx.download("DPDXI")) ^ fingerprint(ctx.download("SU")) ^ fingerprint(ctx.download("P"))
                else:
                    out["assemble"] = fingerprint(ctx.download("A")) ^ fingerprint(ctx.download("FLMASS"))
                    out["gradp_fused"] = fingerprint(ctx.download("U")) ^ fingerprint(ctx.download("W")) ^ fingerprint(ctx.download("P"))
        return out

    variants = []
    if args.round == 1:      # occupancy / L2 prefetch / compact lists / faces per assembly round (profiles/r02_face_ab.txt)
        for cl in (0, 1):
            for occ, pf in ((2, 0), (2, 1), (2, 2), (3, 0), (3, 1), (3, 2)):
                variants.append(dict(FCP_FACE_OCC=occ, FCP_FACE_PF=pf, FCP_FACE_CL=cl, FCP_FACE_AOS=0, FCP_ASM_W=2))
        for pf in (0, 1, 2):
            variants.append(dict(FCP_FACE_OCC=2, FCP_FACE_PF=pf, FCP_FACE_CL=0, FCP_FACE_AOS=0, FCP_ASM_W=1))
    elif args.round == 2:    # array-of-structures face geometry x compact lists (profiles/r02_face_ab2.txt; FCP_FACE_AOS was removed from the library afterwards)
        for aos in (0, 1, 2):
            for cl in (0, 1):
                variants.append(dict(FCP_FACE_OCC=2, FCP_FACE_PF=0, FCP_FACE_CL=cl, FCP_FACE_AOS=aos, FCP_ASM_W=2))
    else:                    # owner-ordered face geometry x compact lists (profiles/r02_face_ab3.txt)
        for og in (0, 1):
            for cl in (0, 1):
                variants.append(dict(FCP_FACE_OCC=2, FCP_FACE_PF=0, FCP_FACE_CL=cl, FCP_FACE_AOS=0, FCP_FACE_OG=og, FCP_ASM_W=2))
    ops = ("grad_gauss", "grad_lsq", "fvx", "gradp_plain", "gradp_fused", "assemble")
    # which switches each operation listens to (the others only repeat a measurement)
    listens = dict(grad_gauss=("FCP_FACE_OCC", "FCP_FACE_PF", "FCP_FACE_CL", "FCP_FACE_AOS", "FCP_FACE_OG"), grad_lsq=("FCP_FACE_OCC", "FCP_FACE_PF", "FCP_FACE_CL"), fvx=("FCP_FACE_CL", "FCP_FACE_OG"),
                   gradp_plain=("FCP_FACE_PF", "FCP_FACE_CL", "FCP_FACE_AOS", "FCP_FACE_OG"), gradp_fused=("FCP_FACE_PF", "FCP_FACE_CL", "FCP_FACE_AOS", "FCP_FACE_OG"),
                   assemble=("FCP_FACE_PF", "FCP_FACE_AOS", "FCP_ASM_W"))
    lines, ref_fp, table = [], None, []
    hdr = f"# face-kernel A/B, {n}^3 hex cavity ({N} cells), ms per launch (mean of {args.reps}), (fraction of the {peak:.0f} GB/s HBM peak); setup {setup_s:.1f} s"
    lines.append(hdr)
    lines.append("# OCC PF CL AOS OG ASM_W | " + " | ".join(f"{o:>20s}" for o in ops) + " | bits")
    print(hdr, flush=True)
    for v in variants:
        v.setdefault("FCP_FACE_OG", 0)
        for k, val in v.items():
            os.environ[k] = str(val)
        fp = run_ops("fp")
        if ref_fp is None:
            ref_fp = fp
        same = all(fp[o] == ref_fp[o] for o in ops)
        ms = run_ops("ms") if same else {o: float("nan") for o in ops}
        table.append((v, ms, same))
        row = (f"  {v['FCP_FACE_OCC']:3d} {v['FCP_FACE_PF']:2d} {v['FCP_FACE_CL']:2d} {v['FCP_FACE_AOS']:3d} {v.get('FCP_FACE_OG', 0):2d} {v['FCP_ASM_W']:5d} | " +
               " | ".join(f"{ms[o]:9.3f} ({nbytes[o] / (ms[o] * 1e-3) / 1e9 / peak:5.3f})   " for o in ops) + (" | same" if same else " | DIFFERENT: " +
                                                                                                              ",".join(o for o in ops if fp[o] != ref_fp[o])))
        lines.append(row)
        print(row, flush=True)
    # best variant per operation, then one environment: OCC / PF / CL from the gradient kernels' and gradp's sum, ASM_W / PF of the assembly reported apart
    best = {}
    for o in ops:
        cand = [(ms[o], v) for v, ms, same in table if same]
        best[o] = min(cand, key=lambda t: t[0])
        lines.append(f"# best {o:12s}: {best[o][0]:.3f} ms  " + " ".join(f"{k}={best[o][1][k]}" for k in listens[o]))

    # per-kernel choice in the order of the library's enum (grad_gauss, grad_lsq, gradp, assemble); gradp: the sum of its two launches
    def best_of(score):
        cand = [(score(ms), v) for v, ms, same in table if same]
        return min(cand, key=lambda t: t[0])[1]
    sel = [best_of(lambda ms: ms["grad_gauss"]), best_of(lambda ms: ms["grad_lsq"]), best_of(lambda ms: ms["gradp_plain"] + ms["gradp_fused"]),
           best_of(lambda ms: ms["assemble"])]
    export = ("export " + " ".join(f"{k}={','.join(str(v[k]) for v in sel)}" for k in ("FCP_FACE_OCC", "FCP_FACE_PF", "FCP_FACE_CL", "FCP_FACE_OG")) +
              f" FCP_ASM_W={sel[3]['FCP_ASM_W']}")
    lines.append("# per kernel (grad_gauss, grad_lsq, gradp, assemble):")
    lines.append(export)
    print("\n".join(lines[-8:]), flush=True)
    if args.out:
        with open(args.out, "w") as fh:
            fh.write("\n".join(lines) + "\n")
        with open(os.path.splitext(args.out)[0] + ".json", "w") as fh:
            json.dump([dict(variant=v, ms=ms, same=same) for v, ms, same in table], fh)
    ctx.close()


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Turns bench.py / bench_rows.py logs (gpurun_out/*.log, scratch) into the committed text tables under profiles/ (development helper).
    python tools/profiles_table.py bench  LABEL=path.log [LABEL=path.log ...]   one row per bench.py line: value, per-iteration time, kernel table
    python tools/profiles_table.py rows   path.log                              tools/bench_rows.py output as a table
"""
import json
import sys


def lines(path):
    out = []
    try:
        for ln in open(path, errors="ignore"):
            ln = ln.strip()
            if ln.startswith("{"):
                try:
                    out.append(json.loads(ln))
                except ValueError:
                    pass
            elif ln.startswith("bench.py:"):
                out.append({"error": ln})
    except OSError as ex:
        out.append({"error": str(ex)})
    return out


def bench(argv):
    keys = ("spmv_dot", "cg_pk", "cg_update", "precond", "assemble", "gradp", "grad_gauss", "grad_lsq")
    print(f"{'run':34s} {'N':>2s} {'value_ms':>9s} {'e2e_ms':>9s} {'iters':>6s} {'us/iter':>8s}  " + " ".join(f"{k + '_us(frac)':>17s}" for k in keys))
    for a in argv:
        label, path = a.split("=", 1)
        for d in lines(path):
            if "error" in d:
                print(f"{label:34s} -- {d['error'][:150]}")
                continue
            if "value" not in d:
                continue
            it = (d.get("pcg") or d.get("iccg") or {}).get("iters") or d.get("config", {}).get("pcg_iters") or 0
            ks = d.get("kernels") or {}
            cells = " ".join((f"{1e3 * ks[k]['avg_ms']:9.1f} ({ks[k]['frac']:.3f})" if k in ks else f"{'-':>17s}") for k in keys)
            print(f"{label:34s} {d.get('n_gpus', 0):2d} {d['value']:9.2f} {d.get('e2e', {}).get('value', 0):9.2f} {it:6d} {1e3 * d['value'] / max(it, 1):8.1f}  {cells}")


def rows(argv):
    print(f"{'operation':58s} {'call_ms':>9s} {'kernel_ms':>9s} {'GB/s':>7s} {'frac':>6s}")
    for d in lines(argv[0]):
        if "op" in d:
            print(f"{d['op'][:58]:58s} {d['call_ms']:9.3f} {d.get('kernel_ms_per_call', 0):9.3f} {d['achieved_gbs']:7.0f} {d['frac']:6.3f}")


if __name__ == "__main__":
    {"bench": bench, "rows": rows}[sys.argv[1]](sys.argv[2:])

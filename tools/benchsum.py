"""Compact one-screen summary of a bench.py JSON line read from stdin (development helper)."""
import json
import sys

KEEP = int(sys.argv[1]) if len(sys.argv) > 1 else 99

for line in sys.stdin:
    line = line.strip()
    if not line.startswith("{"):
        continue
    d = json.loads(line)
    if "unavailable" in d:
        print(d)
        continue
    p = d.get("pcg") or {}
    print(f"n_gpus={d['n_gpus']} value={d['value']:.2f} {d['unit']} e2e={d['e2e']['value']:.2f} iters={p.get('iters')} "
          f"iter_ms={p.get('iteration_ms', 0):.4f} iter_frac={p.get('iteration_frac', 0):.3f} launches={d.get('gpu_launches')}")
    for k, v in list((d.get("kernels") or {}).items())[:KEEP]:
        print(f"   {k:14s} avg_ms={v['avg_ms']:.4f} GB/s={v['achieved_gbs']:.0f} frac={v['frac']:.3f} n={v['launches']}")
    if d.get("cpu_baseline"):
        print("   cpu_baseline", d["cpu_baseline"]["value"], d["cpu_baseline"]["unit"])
    if d.get("clocks") and KEEP > 10:
        print("   clocks", d["clocks"])

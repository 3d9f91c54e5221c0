#!/bin/bash
# Round 2, final one-GPU call: the whole -m gpu suite on the final tree, the contract benchmark and its reference arm, the IC(0)-CG step with and
# without tile prefetch, the polyhedral workload, an ncu capture of the final sweep kernel.
TAG=${1:-r02}
OUT=gpurun_out
mkdir -p $OUT
export PYTHONUNBUFFERED=1
B="python bench.py --no-cpu-baseline"
timeout 1500 python -m pytest tests -m gpu -x -q --durations=5 > $OUT/${TAG}_s8_pytest.log 2>&1
echo "pytest exit $?" >> $OUT/${TAG}_s8_status.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_s8_smoke.log 2>&1
echo "smoke exit $?" >> $OUT/${TAG}_s8_status.txt
timeout 900 python bench.py --steps 5 --warmup 3 > $OUT/${TAG}_s8_bench1.log 2>&1
echo "bench exit $?" >> $OUT/${TAG}_s8_status.txt
timeout 600 $B --solver iccg --steps 2 --warmup 1 > $OUT/${TAG}_s8_iccg_pf.log 2>&1
FCP_SWEEP_PF=off timeout 600 $B --solver iccg --steps 2 --warmup 1 > $OUT/${TAG}_s8_iccg_nopf.log 2>&1
echo "iccg done" >> $OUT/${TAG}_s8_status.txt
timeout 1200 python bench.py --workload poly --steps 3 --warmup 2 > $OUT/${TAG}_s8_poly_n1.log 2>&1
echo "poly exit $?" >> $OUT/${TAG}_s8_status.txt
NCU="ncu --clock-control none"
timeout 600 $NCU --set full --import-source on -k "regex:k_precond_apply_ll" -c 2 -o $OUT/${TAG}_s8_ncu_iccg_ll -f $B --cells 128 --solver iccg --steps 1 --warmup 0 > $OUT/${TAG}_s8_ncu_iccg_ll.log 2>&1
echo "ncu done" >> $OUT/${TAG}_s8_status.txt

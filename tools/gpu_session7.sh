#!/bin/bash
# Round 2, seventh GPU call (one B200): LL sweeps that poll all pending words of a row in one round trip.
TAG=${1:-r02}
OUT=gpurun_out
mkdir -p $OUT
export PYTHONUNBUFFERED=1
B="python bench.py --no-cpu-baseline"
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_sweep_flags.py tests/test_gpu_host_api.py tests/test_gpu_rows2.py -m gpu -x -q -k "csrsolve or sweep or wall_distance or calcp_simple or piso or calcuvw" > $OUT/${TAG}_s7_pytest.log 2>&1
echo "pytest exit $?" >> $OUT/${TAG}_s7_status.txt
timeout 600 $B --solver iccg --steps 2 --warmup 1 > $OUT/${TAG}_s7_iccg_ll.log 2>&1
FCP_SWEEP_CTAS=2 timeout 600 $B --solver iccg --steps 2 --warmup 1 > $OUT/${TAG}_s7_iccg_ll_ctas2.log 2>&1
timeout 600 $B --cells 128 --solver iccg --steps 2 --warmup 1 > $OUT/${TAG}_s7_iccg_ll_n128.log 2>&1
echo "iccg done" >> $OUT/${TAG}_s7_status.txt
NCU="ncu --clock-control none"
timeout 900 $NCU --set full --import-source on -k "regex:k_precond_apply_ll" -c 2 -o $OUT/${TAG}_s7_ncu_iccg_ll -f $B --cells 128 --solver iccg --steps 1 --warmup 0 > $OUT/${TAG}_s7_ncu_iccg_ll.log 2>&1
echo "ncu done" >> $OUT/${TAG}_s7_status.txt

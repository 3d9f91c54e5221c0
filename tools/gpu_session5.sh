#!/bin/bash
# Round 2, fifth GPU call (one B200): programmatic dependent launch A/B, LL sweeps with a sentinel that polls back to back, parity of both on hardware.
TAG=${1:-r02}
OUT=gpurun_out
mkdir -p $OUT
export PYTHONUNBUFFERED=1
B="python bench.py --no-cpu-baseline"
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_zz_simple_loop.py tests/test_gpu_sweep_flags.py -m gpu -x -q > $OUT/${TAG}_s5_pytest.log 2>&1
echo "pytest exit $?" >> $OUT/${TAG}_s5_status.txt
for N in 128 161 256; do
  FCP_PDL=off timeout 600 $B --cells $N --steps 5 --warmup 3 > $OUT/${TAG}_s5_dpcg_n${N}_pdloff.log 2>&1
  timeout 600 $B --cells $N --steps 5 --warmup 3 > $OUT/${TAG}_s5_dpcg_n${N}_pdlon.log 2>&1
done
echo "pdl done" >> $OUT/${TAG}_s5_status.txt
timeout 600 $B --solver iccg --steps 2 --warmup 1 > $OUT/${TAG}_s5_iccg_ll_ns0.log 2>&1
FCP_SWEEP_NS=40 timeout 600 $B --solver iccg --steps 2 --warmup 1 > $OUT/${TAG}_s5_iccg_ll_ns40.log 2>&1
FCP_SWEEP_CTAS=2 timeout 600 $B --solver iccg --steps 2 --warmup 1 > $OUT/${TAG}_s5_iccg_ll_ns0_ctas2.log 2>&1
echo "iccg done" >> $OUT/${TAG}_s5_status.txt
NCU="ncu --clock-control none"
timeout 900 $NCU --set full --import-source on -k "regex:k_precond_apply_ll" -c 2 -o $OUT/${TAG}_s5_ncu_iccg_ll -f $B --cells 128 --solver iccg --steps 1 --warmup 0 > $OUT/${TAG}_s5_ncu_iccg_ll.log 2>&1
echo "ncu done" >> $OUT/${TAG}_s5_status.txt

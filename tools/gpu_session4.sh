#!/bin/bash
# Round 2, fourth GPU call (one B200): staged face kernels, sentinel LL sweeps, L2 carve-out experiment, polyhedral workload again.
TAG=${1:-r02}
OUT=gpurun_out
mkdir -p $OUT
export PYTHONUNBUFFERED=1
B="python bench.py --no-cpu-baseline"
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_rows2.py tests/test_gpu_host_api.py tests/test_gpu_sweep_flags.py -m gpu -x -q > $OUT/${TAG}_s4_pytest.log 2>&1
echo "pytest exit $?" >> $OUT/${TAG}_s4_status.txt
timeout 600 $B --steps 5 --warmup 3 > $OUT/${TAG}_s4_dpcg_n256.log 2>&1
timeout 600 $B --solver iccg --steps 2 --warmup 1 > $OUT/${TAG}_s4_iccg_ll.log 2>&1
FCP_SWEEP_CTAS=2 timeout 600 $B --solver iccg --steps 2 --warmup 1 > $OUT/${TAG}_s4_iccg_ll_ctas2.log 2>&1
FCP_SWEEP_CTAS=1 timeout 600 $B --solver iccg --steps 2 --warmup 1 > $OUT/${TAG}_s4_iccg_ll_ctas1.log 2>&1
echo "iccg done" >> $OUT/${TAG}_s4_status.txt
FCP_L2=carve timeout 600 $B --cells 128 --steps 5 --warmup 3 > $OUT/${TAG}_s4_dpcg_n128_l2carve.log 2>&1
FCP_L2=carve FCP_L2_MB=64 timeout 600 $B --cells 128 --steps 5 --warmup 3 > $OUT/${TAG}_s4_dpcg_n128_l2carve64.log 2>&1
timeout 600 $B --cells 64 --steps 5 --warmup 3 > $OUT/${TAG}_s4_dpcg_n64.log 2>&1
echo "l2 done" >> $OUT/${TAG}_s4_status.txt
timeout 900 python tools/bench_rows.py --n 256 --reps 5 > $OUT/${TAG}_s4_rows.log 2>&1
timeout 1200 python bench.py --workload poly --steps 3 --warmup 2 > $OUT/${TAG}_s4_poly_n1.log 2>&1
echo "poly exit $?" >> $OUT/${TAG}_s4_status.txt
NCU="ncu --clock-control none"
timeout 900 $NCU --set full --import-source on -k "regex:k_gradp|k_assemble_pcorr|k_grad_gauss|k_grad_lsq" -c 6 -o $OUT/${TAG}_s4_ncu_fvm -f python tools/bench_rows.py --n 256 --reps 1 > $OUT/${TAG}_s4_ncu_fvm.log 2>&1
timeout 900 $NCU --set full --import-source on -k "regex:k_precond_apply_ll" -c 2 -o $OUT/${TAG}_s4_ncu_iccg_ll -f $B --cells 128 --solver iccg --steps 1 --warmup 0 > $OUT/${TAG}_s4_ncu_iccg_ll.log 2>&1
echo "ncu done" >> $OUT/${TAG}_s4_status.txt

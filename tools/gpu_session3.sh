#!/bin/bash
# Round 2, third GPU call: the flag-in-data IC(0)/ILU(0) sweeps, the L2 residency hints, the pipelined convergence polling and the polyhedral workload
# on one B200.  Every step under its own timeout, results in gpurun_out/.
TAG=${1:-r02}
OUT=gpurun_out
mkdir -p $OUT
export PYTHONUNBUFFERED=1
B="python bench.py --no-cpu-baseline"
timeout 1500 python -m pytest tests -m gpu -x -q --durations=8 > $OUT/${TAG}_s3_pytest.log 2>&1
echo "pytest exit $?" >> $OUT/${TAG}_s3_status.txt
# IC(0)-CG at 256^3: flag-in-data sweeps (default) with the occupancy-maximal grid and with 1, 2, 4 CTAs per SM
timeout 600 $B --solver iccg --steps 2 --warmup 1 > $OUT/${TAG}_s3_iccg_ll.log 2>&1
for C in 1 2 4; do FCP_SWEEP_CTAS=$C timeout 600 $B --solver iccg --steps 2 --warmup 1 > $OUT/${TAG}_s3_iccg_ll_ctas$C.log 2>&1; done
echo "iccg done" >> $OUT/${TAG}_s3_status.txt
# DPCG: pipelined polling at the contract size; L2 hints off / auto at the per-rank sizes of the 8- and 4-GPU runs (128^3, 161^3) and at 2 GPUs (203^3)
timeout 600 $B --steps 5 --warmup 3 > $OUT/${TAG}_s3_dpcg_n256.log 2>&1
for N in 128 161 203; do
  FCP_L2=off timeout 600 $B --cells $N --steps 5 --warmup 3 > $OUT/${TAG}_s3_dpcg_n${N}_l2off.log 2>&1
  timeout 600 $B --cells $N --steps 5 --warmup 3 > $OUT/${TAG}_s3_dpcg_n${N}_l2auto.log 2>&1
done
FCP_L2=all timeout 600 $B --cells 161 --steps 5 --warmup 3 > $OUT/${TAG}_s3_dpcg_n161_l2all.log 2>&1
FCP_DPCG=persist timeout 600 $B --cells 128 --steps 5 --warmup 3 > $OUT/${TAG}_s3_dpcg_n128_persist3.log 2>&1
FCP_DPCG=persist timeout 600 $B --steps 5 --warmup 3 > $OUT/${TAG}_s3_dpcg_n256_persist3.log 2>&1
echo "dpcg done" >> $OUT/${TAG}_s3_status.txt
# config 5: ~20 M polyhedra on one GPU
timeout 1200 python bench.py --workload poly --steps 3 --warmup 2 > $OUT/${TAG}_s3_poly_n1.log 2>&1
echo "poly exit $?" >> $OUT/${TAG}_s3_status.txt
NCU="ncu --clock-control none"
timeout 900 $NCU --set full --import-source on -k "regex:k_precond_apply_ll|k_factor_ll" -c 3 -o $OUT/${TAG}_s3_ncu_iccg_ll -f $B --cells 128 --solver iccg --steps 1 --warmup 0 > $OUT/${TAG}_s3_ncu_iccg_ll.log 2>&1
timeout 900 $NCU --set full --import-source on -k "regex:k_cg_pk_l2|k_spmv_dot_pipe_l2|k_cg_update_l2" -s 60 -c 3 -o $OUT/${TAG}_s3_ncu_pcg_l2 -f $B --cells 128 --steps 1 --warmup 0 > $OUT/${TAG}_s3_ncu_pcg_l2.log 2>&1
echo "ncu done" >> $OUT/${TAG}_s3_status.txt

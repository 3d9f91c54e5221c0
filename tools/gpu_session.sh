#!/bin/bash
# One gpurun call that collects everything round 1 left unmeasured (profiles/README.md "Not profiled"):
#   /usr/local/graft/bin/gpurun --timeout 2400 -- 'bash tools/gpu_session.sh r02'
# Every step runs under its own `timeout`, writes into gpurun_out/ and never stops the following ones.  Read the results here with
#   python tools/benchsum.py < gpurun_out/<tag>_bench1.log
#   python profiles/summarize.py launches gpurun_out/<tag>_launches.csv profiles/<tag>_launches.txt
#   python profiles/summarize.py ncu      gpurun_out/<tag>_ncu_pcg.ncu-rep profiles/<tag>_ncu_pcg.txt      (same for _fvm, _rows)
# Multi-GPU (separate calls, `gpurun --gpus N`):  bash tools/gpu_session.sh r02 scale N
TAG=${1:-r02}
MODE=${2:-single}
OUT=gpurun_out
mkdir -p $OUT
export PYTHONUNBUFFERED=1
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,power.limit --format=csv > $OUT/${TAG}_gpu.txt 2>&1

if [ "$MODE" = "scale" ]; then
  N=${3:-2}
  for COMM in p2p nccl; do
    FCP_COMM=$COMM timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600 + N)) \
      bench.py --gpus $N --steps 3 --warmup 3 > $OUT/${TAG}_bench${N}_${COMM}.log 2>&1
    echo "bench N=$N $COMM exit $?" >> $OUT/${TAG}_status.txt
  done
  timeout 1500 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > $OUT/${TAG}_pytest_multi${N}.log 2>&1
  echo "pytest multi exit $?" >> $OUT/${TAG}_status.txt
  exit 0
fi

# 1. the parity suite (what the driver runs at round end), with the list of slowest tests
timeout 2400 python -m pytest tests -m gpu -x -q --durations=15 > $OUT/${TAG}_pytest.log 2>&1
echo "pytest exit $?" >> $OUT/${TAG}_status.txt
# 2. smoke + the headline bench line and the reference arm
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1
echo "smoke exit $?" >> $OUT/${TAG}_status.txt
timeout 900 python bench.py --steps 3 --warmup 3 > $OUT/${TAG}_bench1.log 2>&1
echo "bench exit $?" >> $OUT/${TAG}_status.txt
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/${TAG}_ref.log 2>&1
echo "reference arm exit $?" >> $OUT/${TAG}_status.txt
# 2a. the per-rank problem size of the 8-GPU run (256^3 / 8 = 128^3 cells) on ONE GPU: what the three DPCG kernels cost at that size without any
#     communication -- the fixed per-kernel cost that limits strong scaling shows here
timeout 600 python bench.py --cells 128 --steps 5 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_bench1_n128.log 2>&1
echo "bench n=128 exit $?" >> $OUT/${TAG}_status.txt
# 2b. IC(0)/ILU(0): parity of the barrier-free sweeps (opt-in path, never run on hardware in round 1), then A/B timing on the 256^3 ICCG solve, then BiCGStab
FCP_TEST_SWEEP_FLAGS=1 timeout 600 python -m pytest tests/test_gpu_sweep_flags.py -m gpu -x -q > $OUT/${TAG}_pytest_sweep_flags.log 2>&1
echo "sweep-flags parity exit $?" >> $OUT/${TAG}_status.txt
timeout 900 python bench.py --solver iccg --steps 2 --warmup 1 --no-cpu-baseline > $OUT/${TAG}_bench1_iccg_barrier.log 2>&1
FCP_SWEEP=flags timeout 900 python bench.py --solver iccg --steps 2 --warmup 1 --no-cpu-baseline > $OUT/${TAG}_bench1_iccg_flags.log 2>&1
timeout 900 python bench.py --solver bicgstab --steps 2 --warmup 1 --no-cpu-baseline > $OUT/${TAG}_bench1_bicgstab_barrier.log 2>&1
echo "iccg/bicgstab A/B exit $?" >> $OUT/${TAG}_status.txt
# 2c. SpMV variants A/B (default = pipe)
FCP_SPMV=ldg timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $OUT/${TAG}_bench1_spmv_ldg.log 2>&1
FCP_SPMV=tma timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $OUT/${TAG}_bench1_spmv_tma.log 2>&1
echo "spmv A/B exit $?" >> $OUT/${TAG}_status.txt
# 3. per-kernel timings of the rows beyond the headline step (f1, a9, f3, f4): closed cavity and periodic channel
timeout 900 python tools/bench_rows.py --n 256 --reps 5 > $OUT/${TAG}_rows.log 2>&1
timeout 900 python tools/bench_rows.py --n 256 --reps 5 --periodic > $OUT/${TAG}_rows_periodic.log 2>&1
echo "bench_rows exit $?" >> $OUT/${TAG}_status.txt
# 4. ncu: launch list of one step (shares), then full captures of the Krylov kernels, the face kernels and the f-row kernels
NCU="ncu --clock-control none"
timeout 900 $NCU --metrics gpu__time_duration.sum -c 700 --csv --log-file $OUT/${TAG}_launches.csv python bench.py --steps 1 --warmup 0 --prof-steps 0 --no-cpu-baseline > $OUT/${TAG}_ncu_launches.log 2>&1
timeout 900 $NCU --set full --import-source on -k "regex:k_cg_pk|k_spmv_dot|k_cg_update" -s 30 -c 3 -o $OUT/${TAG}_ncu_pcg -f python bench.py --steps 1 --warmup 0 --prof-steps 0 --no-cpu-baseline > $OUT/${TAG}_ncu_pcg.log 2>&1
timeout 900 $NCU --set full --import-source on -k "regex:k_gradp|k_assemble_pcorr|k_correct_flux" -c 4 -o $OUT/${TAG}_ncu_fvm -f python bench.py --steps 1 --warmup 0 --prof-steps 0 --no-cpu-baseline > $OUT/${TAG}_ncu_fvm.log 2>&1
timeout 900 $NCU --set full --import-source on -k "regex:k_precond_apply|k_factor_diag" -c 3 -o $OUT/${TAG}_ncu_iccg -f python bench.py --cells 128 --solver iccg --steps 1 --warmup 0 --prof-steps 0 --no-cpu-baseline > $OUT/${TAG}_ncu_iccg.log 2>&1
timeout 1200 $NCU --set full --import-source on -k "regex:k_uvw_assemble|k_sc_assemble|k_grad_gauss|k_grad_lsq|k_sgs_viscosity|k_piso_H" -c 10 -o $OUT/${TAG}_ncu_rows -f python tools/bench_rows.py --n 128 --reps 1 > $OUT/${TAG}_ncu_rows.log 2>&1
echo "ncu done" >> $OUT/${TAG}_status.txt
ls -la $OUT >> $OUT/${TAG}_status.txt

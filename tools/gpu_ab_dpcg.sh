#!/bin/bash
# A/B of the DPCG drivers on one GPU: three-kernel iteration (FCP_DPCG=kernels) vs the persistent cooperative kernel with 2/3/4 CTAs per SM,
# at the contract size (256^3) and at the per-rank size of the 8-GPU run (128^3).  Parity first (the DPCG tests of the suite run the persistent path).
TAG=${1:-r02}
OUT=gpurun_out
mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_zz_simple_loop.py -m gpu -x -q -k "csrsolve or calcp_simple or determin or simple" > $OUT/${TAG}_ab_pytest.log 2>&1
echo "pytest (persistent DPCG) exit $?" >> $OUT/${TAG}_ab_status.txt
for N in 256 128; do
  FCP_DPCG=kernels timeout 600 python bench.py --cells $N --steps 5 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_ab_n${N}_kernels.log 2>&1
  for MB in 2 3 4; do
    FCP_PERSIST_MINB=$MB timeout 600 python bench.py --cells $N --steps 5 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_ab_n${N}_persist${MB}.log 2>&1
  done
done
echo "A/B done" >> $OUT/${TAG}_ab_status.txt

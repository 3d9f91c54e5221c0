#!/bin/bash
# Round 2, ninth one-GPU call: parity of the code written after call 8 (src-par switches, staged k_uvw_assemble) and the per-row timings again.
TAG=${1:-r02}
OUT=gpurun_out
mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests/test_gpu_rows2.py tests/test_gpu_scalar.py tests/test_gpu_host_api.py tests/test_gpu_zz_simple_loop.py tests/test_gpu_zz_les_channel_loop.py -m gpu -x -q > $OUT/${TAG}_s9_pytest.log 2>&1
echo "pytest exit $?" >> $OUT/${TAG}_s9_status.txt
timeout 900 python tools/bench_rows.py --n 256 --reps 5 > $OUT/${TAG}_s9_rows.log 2>&1
echo "rows exit $?" >> $OUT/${TAG}_s9_status.txt

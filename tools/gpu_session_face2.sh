#!/bin/bash
# Round 2, the very last GPU seconds (~3 minutes of box time left): A/B of the array-of-structures face geometry x compact lists, then the parity
# files and the contract benchmark WITH the switch values that A/B selected.  Every step has its own timeout and writes its results as it goes.
TAG=${1:-r02}
OUT=gpurun_out
mkdir -p $OUT
export PYTHONUNBUFFERED=1
T0=$(date +%s)
stamp() { echo "$1 exit $2 at $(( $(date +%s) - T0 )) s" >> $OUT/${TAG}_face2_status.txt; }
rm -f $OUT/${TAG}_face2_status.txt

timeout 70 python tools/face_ab.py --round 2 --reps 5 --out $OUT/${TAG}_face_ab2.txt > $OUT/${TAG}_face_ab2.log 2>&1
stamp face_ab $?
if grep -q '^export ' $OUT/${TAG}_face_ab2.txt 2>/dev/null; then
  eval "$(grep '^export ' $OUT/${TAG}_face_ab2.txt | tail -1)"
fi
env | grep '^FCP_' > $OUT/${TAG}_face2_env.txt

timeout 75 python -m pytest tests/test_gpu_parity.py tests/test_gpu_rows2.py tests/test_gpu_scalar.py tests/test_gpu_host_api.py tests/test_gpu_gauss_seidel.py -m gpu -x -q > $OUT/${TAG}_face2_pytest.log 2>&1
stamp pytest $?

timeout 60 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_face2_bench1.log 2>&1
stamp bench $?

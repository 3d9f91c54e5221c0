#!/bin/bash
# Round 2, the last GPU seconds (~100 s of box time left): A/B of the owner-ordered face geometry x compact lists, then the parity files with the
# switch values that A/B selected.  Every step has its own timeout and writes its results as it goes.
TAG=${1:-r02}
OUT=gpurun_out
mkdir -p $OUT
export PYTHONUNBUFFERED=1
T0=$(date +%s)
stamp() { echo "$1 exit $2 at $(( $(date +%s) - T0 )) s" >> $OUT/${TAG}_face3_status.txt; }
rm -f $OUT/${TAG}_face3_status.txt

timeout 45 python tools/face_ab.py --round 3 --reps 5 --out $OUT/${TAG}_face_ab3.txt > $OUT/${TAG}_face_ab3.log 2>&1
stamp face_ab $?
if grep -q '^export ' $OUT/${TAG}_face_ab3.txt 2>/dev/null; then
  eval "$(grep '^export ' $OUT/${TAG}_face_ab3.txt | tail -1)"
fi
env | grep '^FCP_' > $OUT/${TAG}_face3_env.txt

timeout 45 python -m pytest tests/test_gpu_parity.py tests/test_gpu_rows2.py tests/test_gpu_scalar.py tests/test_gpu_host_api.py tests/test_gpu_gauss_seidel.py -m gpu -x -q > $OUT/${TAG}_face3_pytest.log 2>&1
stamp pytest $?

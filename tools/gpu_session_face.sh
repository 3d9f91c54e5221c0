#!/bin/bash
# Round 2, the last GPU minutes (one GPU, ~6 minutes of box time): A/B of the face-kernel switches, then the -m gpu suite and the contract
# benchmark WITH the switch values the A/B selected (so that what becomes the default afterwards is what the suite and the bench ran).
# Every step has its own timeout; results are written as they come (the call may be cut by the budget).
TAG=${1:-r02}
OUT=gpurun_out
mkdir -p $OUT
export PYTHONUNBUFFERED=1
T0=$(date +%s)
stamp() { echo "$1 exit $2 at $(( $(date +%s) - T0 )) s" >> $OUT/${TAG}_face_status.txt; }
rm -f $OUT/${TAG}_face_status.txt
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/${TAG}_face_gpu.txt 2>&1

timeout 150 python tools/face_ab.py --reps 5 --out $OUT/${TAG}_face_ab.txt > $OUT/${TAG}_face_ab.log 2>&1
stamp face_ab $?
if grep -q '^export ' $OUT/${TAG}_face_ab.txt 2>/dev/null; then
  eval "$(grep '^export ' $OUT/${TAG}_face_ab.txt | tail -1)"
fi
env | grep '^FCP_' > $OUT/${TAG}_face_env.txt

timeout 200 python -m pytest tests -m gpu -x -q --durations=5 > $OUT/${TAG}_face_pytest.log 2>&1
stamp pytest $?

timeout 120 python tools/bench_rows.py --reps 3 --only "grad_gauss,grad_lsq,limiter,calcsc k (realizable),calcsc omega,calcuvw" > $OUT/${TAG}_face_rows.log 2>&1
stamp rows $?

timeout 150 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_face_bench1.log 2>&1
stamp bench $?

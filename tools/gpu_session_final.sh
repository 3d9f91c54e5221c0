#!/bin/bash
# Round 2, last one-GPU call: the whole -m gpu suite, smoke and the contract benchmark on the FINAL tree.
TAG=${1:-r02}
OUT=gpurun_out
mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout 1200 python -m pytest tests -m gpu -x -q --durations=5 > $OUT/${TAG}_final_pytest.log 2>&1
echo "pytest exit $?" >> $OUT/${TAG}_final_status.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_final_smoke.log 2>&1
echo "smoke exit $?" >> $OUT/${TAG}_final_status.txt
timeout 600 python bench.py --steps 5 --warmup 3 > $OUT/${TAG}_final_bench1.log 2>&1
echo "bench exit $?" >> $OUT/${TAG}_final_status.txt

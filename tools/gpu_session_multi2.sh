#!/bin/bash
# Second `gpurun --gpus 8` call of round 2: the residual-halo scheme of DPCG (default) against the direction-vector push (FCP_HALO=pk) on 8 GPUs,
# after one 8-rank parity case on hardware.
TAG=${1:-r02}
OUT=gpurun_out
mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q -p no:cacheprovider -k "8-p2p-slab" > $OUT/${TAG}_multi2_t8.log 2>&1
echo "8-rank parity exit $?" >> $OUT/${TAG}_multi2_status.txt
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 600 $TR --nproc-per-node 8 --master-port 29621 bench.py --gpus 8 --steps 10 --warmup 3 > $OUT/${TAG}_multi2_bench8_reshalo.log 2>&1
echo "bench8 res-halo exit $?" >> $OUT/${TAG}_multi2_status.txt
FCP_HALO=pk timeout 600 $TR --nproc-per-node 8 --master-port 29622 bench.py --gpus 8 --steps 10 --warmup 3 > $OUT/${TAG}_multi2_bench8_pkhalo.log 2>&1
echo "bench8 pk-halo exit $?" >> $OUT/${TAG}_multi2_status.txt

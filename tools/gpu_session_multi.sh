#!/bin/bash
# One `gpurun --gpus 8` call: (1) tests/test_gpu_multi.py, all cases, on real GPUs -- 2-rank cases four at a time and 4-rank cases two at a time on
# disjoint GPU sets (CUDA_VISIBLE_DEVICES), 8-rank cases one after the other; (2) the contract benchmark on 8 GPUs; (3) the polyhedral workload
# (BASELINE config 5) on 8 GPUs, then on 4 + 2 + 1 GPUs side by side on disjoint GPUs.
#   /usr/local/graft/bin/gpurun --gpus 8 --timeout 1200 -- 'bash tools/gpu_session_multi.sh r02'
TAG=${1:-r02}
OUT=gpurun_out
mkdir -p $OUT
export PYTHONUNBUFFERED=1
nvidia-smi -L > $OUT/${TAG}_multi_gpus.txt 2>&1
PT="python -m pytest tests/test_gpu_multi.py -m gpu -q -p no:cacheprovider"
# ---- (1) parity of the partitioned path
CUDA_VISIBLE_DEVICES=0,1 timeout 900 $PT -k "2-p2p-slab or 2-p2p-poly or 2-nccl-inout" > $OUT/${TAG}_multi_t2a.log 2>&1 &
CUDA_VISIBLE_DEVICES=2,3 timeout 900 $PT -k "2-nccl-slab or 2-nccl-poly or 2-p2p-inout" > $OUT/${TAG}_multi_t2b.log 2>&1 &
CUDA_VISIBLE_DEVICES=4,5 timeout 900 $PT -k "2-p2p-periodic or 2-p2p-brick" > $OUT/${TAG}_multi_t2c.log 2>&1 &
CUDA_VISIBLE_DEVICES=6,7 timeout 900 $PT -k "2-nccl-periodic or 2-nccl-brick" > $OUT/${TAG}_multi_t2d.log 2>&1 &
wait
CUDA_VISIBLE_DEVICES=0,1,2,3 timeout 900 $PT -k "4-p2p" > $OUT/${TAG}_multi_t4a.log 2>&1 &
CUDA_VISIBLE_DEVICES=4,5,6,7 timeout 900 $PT -k "4-nccl" > $OUT/${TAG}_multi_t4b.log 2>&1 &
wait
timeout 900 $PT -k "8-" > $OUT/${TAG}_multi_t8.log 2>&1
tail -n 3 $OUT/${TAG}_multi_t*.log > $OUT/${TAG}_multi_tests_summary.txt 2>&1
echo "multi tests done" >> $OUT/${TAG}_multi_status.txt
# ---- (2) contract benchmark, 8 GPUs: default, then own-window polls at GPU scope
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 600 $TR --nproc-per-node 8 --master-port 29611 bench.py --gpus 8 --steps 10 --warmup 3 > $OUT/${TAG}_multi_bench8.log 2>&1
echo "bench8 exit $?" >> $OUT/${TAG}_multi_status.txt
FCP_P2P_POLL=gpu timeout 600 $TR --nproc-per-node 8 --master-port 29612 bench.py --gpus 8 --steps 10 --warmup 3 > $OUT/${TAG}_multi_bench8_pollgpu.log 2>&1
echo "bench8 poll=gpu exit $?" >> $OUT/${TAG}_multi_status.txt
# ---- (3) polyhedral workload: 8 GPUs, then 4 + 2 GPUs side by side on disjoint GPUs
timeout 900 $TR --nproc-per-node 8 --master-port 29613 bench.py --gpus 8 --workload poly --steps 3 --warmup 2 > $OUT/${TAG}_multi_poly8.log 2>&1
echo "poly8 exit $?" >> $OUT/${TAG}_multi_status.txt
CUDA_VISIBLE_DEVICES=0,1,2,3 timeout 900 $TR --nproc-per-node 4 --master-port 29614 bench.py --gpus 4 --workload poly --steps 3 --warmup 2 > $OUT/${TAG}_multi_poly4.log 2>&1 &
CUDA_VISIBLE_DEVICES=4,5 timeout 900 $TR --nproc-per-node 2 --master-port 29615 bench.py --gpus 2 --workload poly --steps 3 --warmup 2 > $OUT/${TAG}_multi_poly2.log 2>&1 &
CUDA_VISIBLE_DEVICES=6,7 timeout 600 $TR --nproc-per-node 2 --master-port 29617 bench.py --gpus 2 --steps 10 --warmup 3 > $OUT/${TAG}_multi_bench2.log 2>&1 &
wait
echo "poly4/2 + bench2 done" >> $OUT/${TAG}_multi_status.txt

#!/bin/bash
# round-2 GPU session R (8 GPUs): peer-window all-reduce at N = 8 against ncclAllReduce, C4 as specified
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/r_gpus.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29521 bench.py --gpus 8 --steps 4 --warmup 2 --no-cpu-baseline > gpurun_out/r_bench_n8.json 2> gpurun_out/r_bench_n8.err
echo "bench rc=$?" >> gpurun_out/r_bench_n8.err
DIAGLIB_B200_PEER_REDUCE=0 timeout 300 $TR --master-port 29522 bench.py --gpus 8 --steps 4 --warmup 2 --no-cpu-baseline > gpurun_out/r_bench_n8_nccl.json 2> gpurun_out/r_bench_n8_nccl.err
echo "bench-nccl rc=$?" >> gpurun_out/r_bench_n8_nccl.err
timeout 300 $TR --master-port 29523 bench.py --gpus 8 --workload c4 --bits 26 --steps 2 --warmup 1 > gpurun_out/r_c4_n26.json 2> gpurun_out/r_c4_n26.err
echo "c4 rc=$?" >> gpurun_out/r_c4_n26.err
tail -n 2 gpurun_out/r_bench_n8.err gpurun_out/r_bench_n8_nccl.err gpurun_out/r_c4_n26.err

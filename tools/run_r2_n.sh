#!/bin/bash
# round-2 GPU session N (2 GPUs): peer-window all-reduce against ncclAllReduce
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/n_gpus.txt
nvidia-smi topo -m >> gpurun_out/n_gpus.txt 2>&1
timeout 1200 python -m pytest tests/test_gpu_multi.py -m gpu -q --timeout 600 --timeout-method=thread > gpurun_out/n_multi.log 2>&1
echo "multi rc=$?" >> gpurun_out/n_multi.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/n_bench_n2.json 2> gpurun_out/n_bench_n2.err
echo "bench rc=$?" >> gpurun_out/n_bench_n2.err
DIAGLIB_B200_PEER_REDUCE=0 timeout 600 $TR --master-port 29512 bench.py --gpus 2 --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/n_bench_n2_nccl.json 2> gpurun_out/n_bench_n2_nccl.err
echo "bench rc=$?" >> gpurun_out/n_bench_n2_nccl.err
tail -n 5 gpurun_out/n_multi.log gpurun_out/n_bench_n2.err gpurun_out/n_bench_n2_nccl.err

#!/bin/bash
# final build on 2 GPUs: two-rank tests, bench line
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q --timeout 600 --timeout-method=thread > gpurun_out/n2_multi.log 2>&1
echo "multi rc=$?" >> gpurun_out/n2_multi.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 400 $TR --master-port 29511 bench.py --gpus 2 --steps 4 --warmup 2 --no-cpu-baseline > gpurun_out/n2_bench.json 2> gpurun_out/n2_bench.err
echo "bench rc=$?" >> gpurun_out/n2_bench.err
tail -n 3 gpurun_out/n2_multi.log gpurun_out/n2_bench.err

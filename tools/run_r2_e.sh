#!/bin/bash
# round-2 GPU session E (8 GPUs): C3 at N=8 (default and with the round-1 control flow), C4 as specified (n = 2^26)
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name,memory.total --format=csv > gpurun_out/e_gpus.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29521 bench.py --gpus 8 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/e_bench_n8.json 2> gpurun_out/e_bench_n8.err
echo "bench rc=$?" >> gpurun_out/e_bench_n8.err
timeout 900 $TR --master-port 29523 bench.py --gpus 8 --workload c4 --bits 26 --steps 2 --warmup 1 > gpurun_out/e_c4_n26.json 2> gpurun_out/e_c4_n26.err
echo "c4 rc=$?" >> gpurun_out/e_c4_n26.err
DIAGLIB_B200_SPEC_ORTHO=0 DIAGLIB_B200_HALO_OVERLAP=0 timeout 600 $TR --master-port 29522 bench.py --gpus 8 --steps 5 --warmup 3 --no-cpu-baseline --natural-order > gpurun_out/e_bench_n8_r1flow.json 2> gpurun_out/e_bench_n8_r1flow.err
echo "bench-r1flow rc=$?" >> gpurun_out/e_bench_n8_r1flow.err
tail -n 3 gpurun_out/e_bench_n8.err gpurun_out/e_c4_n26.err gpurun_out/e_bench_n8_r1flow.err

#!/bin/bash
# round-2 GPU session D (2 GPUs): multi-rank parity tests, halo overlap on/off, C4 on two ranks
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/d_gpus.txt
timeout 1200 python -m pytest tests/test_gpu_multi.py -m gpu -q -s --timeout 600 --timeout-method=thread > gpurun_out/d_multi.log 2>&1
echo "multi rc=$?" >> gpurun_out/d_multi.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 2 > gpurun_out/d_bench_n2.json 2> gpurun_out/d_bench_n2.err
echo "bench rc=$?" >> gpurun_out/d_bench_n2.err
DIAGLIB_B200_HALO_OVERLAP=0 timeout 600 $TR --master-port 29512 bench.py --gpus 2 --steps 3 --warmup 2 > gpurun_out/d_bench_n2_nooverlap.json 2> gpurun_out/d_bench_n2_nooverlap.err
timeout 600 $TR --master-port 29513 bench.py --gpus 2 --workload c4 --bits 23 --steps 2 --warmup 1 > gpurun_out/d_c4_n23.json 2> gpurun_out/d_c4_n23.err
echo "c4 rc=$?" >> gpurun_out/d_c4_n23.err
timeout 600 $TR --master-port 29514 bench.py --gpus 2 --workload c4 --bits 22 --steps 2 --warmup 1 > gpurun_out/d_c4_n22.json 2> gpurun_out/d_c4_n22.err
echo "c4-22 rc=$?" >> gpurun_out/d_c4_n22.err
DIAGLIB_B200_SPMM_TAIL=2 timeout 300 python tools/spmm_order_bench.py 256 tail2 64x2x2 > gpurun_out/d_spmm_tail2.log 2>&1
timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "sym_eig_timing or spmm" -s > gpurun_out/d_kern.log 2>&1
tail -n 4 gpurun_out/d_multi.log gpurun_out/d_bench_n2.err gpurun_out/d_c4_n23.err gpurun_out/d_c4_n22.err

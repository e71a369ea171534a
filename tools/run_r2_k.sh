#!/bin/bash
# round-2 GPU session K (1 GPU): compile-time triangular chunks + direct-add identity chunks
mkdir -p gpurun_out
timeout 300 python tools/kernel_bench.py 24 5 trmm > gpurun_out/k_trmm.json 2> gpurun_out/k_trmm.err
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q --timeout 300 --timeout-method=thread > gpurun_out/k_kernels.log 2>&1
echo "kernels rc=$?" >> gpurun_out/k_kernels.log
timeout 1800 python -m pytest tests/test_gpu_drivers.py tests/test_abi.py -m gpu -q --timeout 600 --timeout-method=thread > gpurun_out/k_drivers.log 2>&1
echo "drivers rc=$?" >> gpurun_out/k_drivers.log
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/k_bench.json 2> gpurun_out/k_bench.err
echo "bench rc=$?" >> gpurun_out/k_bench.err
cat gpurun_out/k_trmm.json; tail -n 3 gpurun_out/k_kernels.log gpurun_out/k_drivers.log gpurun_out/k_bench.err

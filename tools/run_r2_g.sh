#!/bin/bash
# round-2 GPU session G (1 GPU): shared-memory chol_inv / get_coeffs, suites, bench line
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q -s --timeout 300 --timeout-method=thread > gpurun_out/g_kernels.log 2>&1
echo "kernels rc=$?" >> gpurun_out/g_kernels.log
timeout 1800 python -m pytest tests/test_gpu_drivers.py tests/test_abi.py -m gpu -q -s --timeout 600 --timeout-method=thread > gpurun_out/g_drivers.log 2>&1
echo "drivers rc=$?" >> gpurun_out/g_drivers.log
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/g_bench.json 2> gpurun_out/g_bench.err
echo "bench rc=$?" >> gpurun_out/g_bench.err
timeout 600 python tools/c5_run.py 22 c5 > gpurun_out/g_c5.json 2> gpurun_out/g_c5.err
tail -n 3 gpurun_out/g_kernels.log gpurun_out/g_drivers.log gpurun_out/g_bench.err

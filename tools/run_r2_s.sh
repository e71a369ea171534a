#!/bin/bash
# round-2 GPU session S: refined Gram schedules, peer time-out check
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q -s --timeout 300 --timeout-method=thread > gpurun_out/s_kernels.log 2>&1
echo "kernels rc=$?" >> gpurun_out/s_kernels.log
timeout 1800 python -m pytest tests/test_gpu_drivers.py tests/test_abi.py -m gpu -q --timeout 600 --timeout-method=thread > gpurun_out/s_drivers.log 2>&1
echo "drivers rc=$?" >> gpurun_out/s_drivers.log
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/s_bench.json 2> gpurun_out/s_bench.err
echo "bench rc=$?" >> gpurun_out/s_bench.err
grep -h "chol_inv m=\|get_coeffs len_u" gpurun_out/s_kernels.log
tail -n 3 gpurun_out/s_kernels.log gpurun_out/s_drivers.log gpurun_out/s_bench.err

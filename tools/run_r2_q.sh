#!/bin/bash
# round-2 GPU session Q: get_coeffs with staged eigenvectors / ILP, trimmed diagonal block
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q -s --timeout 300 --timeout-method=thread > gpurun_out/q_kernels.log 2>&1
echo "kernels rc=$?" >> gpurun_out/q_kernels.log
timeout 1800 python -m pytest tests/test_gpu_drivers.py tests/test_abi.py -m gpu -q --timeout 600 --timeout-method=thread > gpurun_out/q_drivers.log 2>&1
echo "drivers rc=$?" >> gpurun_out/q_drivers.log
grep -h "chol_inv m=\|get_coeffs len_u" gpurun_out/q_kernels.log
tail -n 3 gpurun_out/q_kernels.log gpurun_out/q_drivers.log

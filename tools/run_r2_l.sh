#!/bin/bash
# round-2 GPU session L (1 GPU): all triangular widths compile-time, projection hook; suites + bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q --timeout 300 --timeout-method=thread > gpurun_out/l_kernels.log 2>&1
echo "kernels rc=$?" >> gpurun_out/l_kernels.log
timeout 1800 python -m pytest tests/test_gpu_drivers.py tests/test_abi.py -m gpu -q --timeout 600 --timeout-method=thread > gpurun_out/l_drivers.log 2>&1
echo "drivers rc=$?" >> gpurun_out/l_drivers.log
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/l_bench.json 2> gpurun_out/l_bench.err
echo "bench rc=$?" >> gpurun_out/l_bench.err
timeout 600 python tools/c5_run.py 22 c5 > gpurun_out/l_c5.json 2> gpurun_out/l_c5.err
timeout 600 python bench.py --workload c4 --bits 22 --steps 2 --warmup 1 > gpurun_out/l_c4_n22.json 2> gpurun_out/l_c4_n22.err
tail -n 3 gpurun_out/l_kernels.log gpurun_out/l_drivers.log gpurun_out/l_bench.err

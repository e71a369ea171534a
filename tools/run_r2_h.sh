#!/bin/bash
# round-2 GPU session H (1 GPU): final build - suites, timing tables, bench line, smoke
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q -s --timeout 300 --timeout-method=thread > gpurun_out/h_kernels.log 2>&1
echo "kernels rc=$?" >> gpurun_out/h_kernels.log
timeout 1800 python -m pytest tests/test_gpu_drivers.py tests/test_abi.py -m gpu -q -s --timeout 600 --timeout-method=thread > gpurun_out/h_drivers.log 2>&1
echo "drivers rc=$?" >> gpurun_out/h_drivers.log
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/h_bench.json 2> gpurun_out/h_bench.err
echo "bench rc=$?" >> gpurun_out/h_bench.err
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/h_smoke.log 2>&1
echo "smoke rc=$?" >> gpurun_out/h_smoke.log
tail -n 3 gpurun_out/h_kernels.log gpurun_out/h_drivers.log gpurun_out/h_bench.err gpurun_out/h_smoke.log

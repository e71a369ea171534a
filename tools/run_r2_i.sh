#!/bin/bash
# round-2 GPU session I (1 GPU): projection step as one product over [x u]; suites + bench A/B
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q --timeout 300 --timeout-method=thread > gpurun_out/i_kernels.log 2>&1
echo "kernels rc=$?" >> gpurun_out/i_kernels.log
timeout 1800 python -m pytest tests/test_gpu_drivers.py tests/test_abi.py -m gpu -q -s --timeout 600 --timeout-method=thread > gpurun_out/i_drivers.log 2>&1
echo "drivers rc=$?" >> gpurun_out/i_drivers.log
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/i_bench.json 2> gpurun_out/i_bench.err
echo "bench rc=$?" >> gpurun_out/i_bench.err
DIAGLIB_B200_NO_IDENT_PROJ=1 timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/i_bench_noident.json 2> gpurun_out/i_bench_noident.err
timeout 600 python tools/c5_run.py 22 c5 > gpurun_out/i_c5.json 2> gpurun_out/i_c5.err
tail -n 3 gpurun_out/i_kernels.log gpurun_out/i_drivers.log gpurun_out/i_bench.err

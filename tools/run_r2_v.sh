#!/bin/bash
# round-2 GPU session V: deferred triangular multiply folded into the projection step (A/B in one session)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q --timeout 300 --timeout-method=thread > gpurun_out/v_kernels.log 2>&1
echo "kernels rc=$?" >> gpurun_out/v_kernels.log
timeout 1800 python -m pytest tests/test_gpu_drivers.py tests/test_abi.py -m gpu -q --timeout 600 --timeout-method=thread > gpurun_out/v_drivers.log 2>&1
echo "drivers rc=$?" >> gpurun_out/v_drivers.log
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/v_bench.json 2> gpurun_out/v_bench.err
echo "bench rc=$?" >> gpurun_out/v_bench.err
DIAGLIB_B200_FOLD_TRMM=0 timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/v_bench_nofold.json 2> gpurun_out/v_bench_nofold.err
echo "bench rc=$?" >> gpurun_out/v_bench_nofold.err
tail -n 5 gpurun_out/v_kernels.log gpurun_out/v_drivers.log gpurun_out/v_bench.err gpurun_out/v_bench_nofold.err

#!/bin/bash
# round-2 GPU session F (1 GPU): full suites with the final build, bench line, ncu launch list, SpMM traffic
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q -s --timeout 300 --timeout-method=thread > gpurun_out/f_kernels.log 2>&1
echo "kernels rc=$?" >> gpurun_out/f_kernels.log
timeout 1800 python -m pytest tests/test_gpu_drivers.py tests/test_abi.py tests/test_gpu_multi.py -m gpu -q -s --timeout 600 --timeout-method=thread > gpurun_out/f_drivers.log 2>&1
echo "drivers rc=$?" >> gpurun_out/f_drivers.log
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/f_bench.json 2> gpurun_out/f_bench.err
echo "bench rc=$?" >> gpurun_out/f_bench.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4200 --csv --log-file gpurun_out/launches_r02.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/f_ncu_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:spmm_csr_short -c 4 -o gpurun_out/prof_spmm_r02 -f python tools/spmm_ncu_target.py > gpurun_out/f_ncu_spmm.log 2>&1
tail -n 3 gpurun_out/f_kernels.log gpurun_out/f_drivers.log gpurun_out/f_bench.err gpurun_out/f_ncu_spmm.log

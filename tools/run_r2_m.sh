#!/bin/bash
# round-2 GPU session M: C5 regression fix, sliced gram_reduce, SpMM remainder variants
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q --timeout 300 --timeout-method=thread > gpurun_out/m_kernels.log 2>&1
echo "kernels rc=$?" >> gpurun_out/m_kernels.log
timeout 1800 python -m pytest tests/test_gpu_drivers.py tests/test_abi.py -m gpu -q --timeout 600 --timeout-method=thread > gpurun_out/m_drivers.log 2>&1
echo "drivers rc=$?" >> gpurun_out/m_drivers.log
timeout 300 python tools/spmm_tail_bench.py 256 > gpurun_out/m_spmm_tail.log 2>&1
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/m_bench.json 2> gpurun_out/m_bench.err
echo "bench rc=$?" >> gpurun_out/m_bench.err
timeout 600 python tools/c5_run.py 22 c5 > gpurun_out/m_c5.json 2> gpurun_out/m_c5.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"gram_reduce|chol_inv|get_coeffs|sym_eig" -c 300 --csv --log-file gpurun_out/m_small_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/m_ncu.log 2>&1
tail -n 3 gpurun_out/m_kernels.log gpurun_out/m_drivers.log gpurun_out/m_bench.err gpurun_out/m_c5.err gpurun_out/m_spmm_tail.log

#!/bin/bash
# round-2 GPU session C
mkdir -p gpurun_out
(python tools/oracle_spread.py 256 16 0 gpurun_out/spread_nx256_acc.json 1 > gpurun_out/spread_acc.log 2>&1) &
P1=$!
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q -s --timeout 300 --timeout-method=thread > gpurun_out/c_kernels.log 2>&1
echo "kernels rc=$?" >> gpurun_out/c_kernels.log
timeout 1800 python -m pytest tests/test_gpu_drivers.py tests/test_abi.py -m gpu -q -s --timeout 600 --timeout-method=thread > gpurun_out/c_drivers.log 2>&1
echo "drivers rc=$?" >> gpurun_out/c_drivers.log
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/c_bench.json 2> gpurun_out/c_bench.err
echo "bench rc=$?" >> gpurun_out/c_bench.err
DIAGLIB_B200_SPEC_ORTHO=0 timeout 900 python bench.py --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/c_bench_nospec.json 2> gpurun_out/c_bench_nospec.err
T="64x2x2 64x4x1 128x2x1 64x2x4 128x1x2"
timeout 300 python tools/spmm_order_bench.py 256 tail1_minb4 $T > gpurun_out/c_spmm_a.log 2>&1
DIAGLIB_B200_SPMM_TAIL=0 timeout 300 python tools/spmm_order_bench.py 256 tail0_minb4 64x2x2 > gpurun_out/c_spmm_b.log 2>&1
DIAGLIB_B200_SPMM_MINB=3 timeout 300 python tools/spmm_order_bench.py 256 tail1_minb3 64x2x2 > gpurun_out/c_spmm_c.log 2>&1
DIAGLIB_B200_SPMM_CHUNK_TILED=1 timeout 300 python tools/spmm_order_bench.py 256 tail1_minb4_chunk 64x2x2 > gpurun_out/c_spmm_d.log 2>&1
timeout 600 python tools/c5_run.py 22 c5 > gpurun_out/c_c5.json 2> gpurun_out/c_c5.err
# ncu: one-sided eigensolver, full set with source
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sym_eig_osj -c 2 -s 2 -o gpurun_out/prof_osj_r02 -f python tools/eig_ncu_target.py > gpurun_out/c_ncu_osj.log 2>&1
wait $P1
tail -n 3 gpurun_out/c_kernels.log gpurun_out/c_drivers.log gpurun_out/c_bench.err gpurun_out/spread_acc.log gpurun_out/c_c5.err

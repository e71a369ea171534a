#!/bin/bash
# round-2 GPU session B
mkdir -p gpurun_out
(python tools/oracle_spread.py 256 8 0 gpurun_out/spread_nx256_t8.json > gpurun_out/spread_t8.log 2>&1) &
P1=$!
(python tools/oracle_spread.py 256 8 1e-14 gpurun_out/spread_nx256_t8_p.json > gpurun_out/spread_t8p.log 2>&1) &
P2=$!
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q -s --timeout 300 --timeout-method=thread > gpurun_out/b_kernels.log 2>&1
echo "kernels rc=$?" >> gpurun_out/b_kernels.log
timeout 1800 python -m pytest tests/test_gpu_drivers.py tests/test_abi.py tests/test_gpu_multi.py -m gpu -q --timeout 600 --timeout-method=thread > gpurun_out/b_drivers.log 2>&1
echo "drivers rc=$?" >> gpurun_out/b_drivers.log
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/b_bench.json 2> gpurun_out/b_bench.err
echo "bench rc=$?" >> gpurun_out/b_bench.err
timeout 600 python tools/spmm_order_bench.py 256 > gpurun_out/b_spmm_order.log 2>&1
timeout 120 tools/dmma_bench > gpurun_out/b_dmma.txt 2>&1
timeout 900 python bench.py --workload c4 --bits 22 --steps 2 --warmup 1 > gpurun_out/b_c4_n22.json 2> gpurun_out/b_c4_n22.err
echo "c4 rc=$?" >> gpurun_out/b_c4_n22.err
wait $P1 $P2
tail -n 3 gpurun_out/b_kernels.log gpurun_out/b_drivers.log gpurun_out/b_bench.err gpurun_out/b_c4_n22.err gpurun_out/spread_t8.log gpurun_out/spread_t8p.log

#!/bin/bash
# round-2 GPU session J (1 GPU): in-place vs out-of-place triangular multiply, ncu of the in-place one
mkdir -p gpurun_out
timeout 300 python tools/kernel_bench.py 24 5 trmm > gpurun_out/j_trmm.json 2> gpurun_out/j_trmm.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:blockmul_ws -s 1 -c 2 -o gpurun_out/prof_trmm_r02 -f python tools/kernel_bench.py 24 1 trmm > gpurun_out/j_ncu_trmm.log 2>&1
cat gpurun_out/j_trmm.json; tail -n 3 gpurun_out/j_ncu_trmm.log

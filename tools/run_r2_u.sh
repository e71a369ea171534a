#!/bin/bash
# round-2 GPU session U: projection kernel without the alpha multiply, 256-row tiles for comparison, ncu of the 74 x 37 Gram
mkdir -p gpurun_out
python tools/kernel_bench.py 24 5 ortho > gpurun_out/u_ortho_n24.json 2> gpurun_out/u_ortho.err
DIAGLIB_B200_BMUL_RT256=1 python tools/kernel_bench.py 24 5 ortho > gpurun_out/u_ortho_n24_rt256.json 2>> gpurun_out/u_ortho.err
python tools/kernel_bench.py 24 5 trmm > gpurun_out/u_trmm_n24.json 2>> gpurun_out/u_ortho.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"gram_tma" -c 1 -s 1 -o gpurun_out/prof_gram7437_r02 -f python tools/kernel_bench.py 22 1 ortho > gpurun_out/u_ncu.log 2>&1
cat gpurun_out/u_ortho_n24.json gpurun_out/u_ortho_n24_rt256.json gpurun_out/u_trmm_n24.json; tail -n 3 gpurun_out/u_ncu.log

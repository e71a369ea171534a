#!/bin/bash
# round-2 GPU session T: ncu --set full with source of the two wide ortho_vs_x kernels (Gram 74 x 37, projection)
mkdir -p gpurun_out
python tools/kernel_bench.py 24 5 ortho > gpurun_out/t_ortho_n24.json 2> gpurun_out/t_ortho.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"gram_tma|blockmul_ws" -c 2 -s 2 -o gpurun_out/prof_ortho_r02 -f python tools/kernel_bench.py 22 1 ortho > gpurun_out/t_ncu.log 2>&1
cat gpurun_out/t_ortho_n24.json; tail -n 4 gpurun_out/t_ncu.log

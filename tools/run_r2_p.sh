#!/bin/bash
# round-2 GPU session P: where get_coeffs spends its time (ncu source view)
mkdir -p gpurun_out
python tools/coeffs_ncu_target.py > gpurun_out/p_coeffs.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:get_coeffs -c 1 -s 1 -o gpurun_out/prof_coeffs_r02 -f python tools/coeffs_ncu_target.py > gpurun_out/p_ncu.log 2>&1
tail -n 5 gpurun_out/p_coeffs.log gpurun_out/p_ncu.log

#!/bin/bash
# SpMM variants back to back on one box: generic (0) vs short-row kernel (1)
for s in ${CFGS:-0 1}; do
  echo "short=$s"
  DIAGLIB_B200_SPMM_SHORT=$s python tools/kernel_bench.py 24 ${REPS:-5} spmm
done

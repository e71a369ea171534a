#!/bin/bash
# SpMM variants back to back on one box (short:chunk): generic kernel, short-row kernel, column chunks
for cfg in ${CFGS:-0:0 1:0 1:16 1:24 1:32}; do
  IFS=: read s c <<< "$cfg"
  echo "short=$s chunk=$c"
  DIAGLIB_B200_SPMM_SHORT=$s DIAGLIB_B200_SPMM_CHUNK=$c python tools/kernel_bench.py 24 ${REPS:-5} spmm
done

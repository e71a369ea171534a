#!/bin/bash
mkdir -p gpurun_out
cp diaglib_b200/libdiaglib_b200.so /tmp/shipped.so
(cp build/dbg_1.so diaglib_b200/libdiaglib_b200.so
 echo "== direct-store variant, chain, fold"; python tools/seq_repro.py 21 25
 echo "== direct-store variant, chain, no fold"; DIAGLIB_B200_FOLD_TRMM=0 python tools/seq_repro.py 21 25
 echo "== direct-store variant, host-driven, no fold"; DIAGLIB_B200_SPEC_ORTHO=0 DIAGLIB_B200_FOLD_TRMM=0 python tools/seq_repro.py 21 25
 cp /tmp/shipped.so diaglib_b200/libdiaglib_b200.so
 echo "== shipped"; python tools/seq_repro.py 21 25) > gpurun_out/dbg3.log 2>&1
cat gpurun_out/dbg3.log

#!/bin/bash
mkdir -p gpurun_out
cp diaglib_b200/libdiaglib_b200.so /tmp/shipped.so
(cp build/dbg_1.so diaglib_b200/libdiaglib_b200.so
 echo "== direct-store variant"; python tools/kernel_repro.py 21 30
 cp /tmp/shipped.so diaglib_b200/libdiaglib_b200.so
 echo "== shipped"; python tools/kernel_repro.py 21 30) > gpurun_out/dbg2.log 2>&1
cat gpurun_out/dbg2.log

#!/bin/bash
mkdir -p gpurun_out
cp diaglib_b200/libdiaglib_b200.so /tmp/shipped.so
cp build/dbg_1.so diaglib_b200/libdiaglib_b200.so
(for d in 0 3 4 8; do echo "== direct-store, DBG=$d"; DIAGLIB_B200_FOLD_TRMM=0 DIAGLIB_B200_DBG=$d python tools/determinism_check.py 128 10; done) > gpurun_out/dbg6.log 2>&1
cp /tmp/shipped.so diaglib_b200/libdiaglib_b200.so
grep "^==\|DETERM" gpurun_out/dbg6.log

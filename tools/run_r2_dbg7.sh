#!/bin/bash
mkdir -p gpurun_out
cp diaglib_b200/libdiaglib_b200.so /tmp/shipped.so
cp build/dbg_1.so diaglib_b200/libdiaglib_b200.so
(echo "== direct-store, one CTA per SM (DBG=16)"; DIAGLIB_B200_FOLD_TRMM=0 DIAGLIB_B200_DBG=16 python tools/determinism_check.py 128 10
 echo "== direct-store, 256-row tiles / 16 consumer warps"; DIAGLIB_B200_FOLD_TRMM=0 DIAGLIB_B200_BMUL_RT256=1 python tools/determinism_check.py 128 10) > gpurun_out/dbg7.log 2>&1
cp /tmp/shipped.so diaglib_b200/libdiaglib_b200.so
grep "^==\|DETERM" gpurun_out/dbg7.log

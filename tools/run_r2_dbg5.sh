#!/bin/bash
mkdir -p gpurun_out
cp diaglib_b200/libdiaglib_b200.so /tmp/shipped.so
cp build/dbg_1.so diaglib_b200/libdiaglib_b200.so
(echo "== direct-store, no ws Gram (mask 1)"; DIAGLIB_B200_FOLD_TRMM=0 DIAGLIB_B200_WS_MASK=1 python tools/determinism_check.py 128 10
 echo "== direct-store, no ws block multiply (mask 2)"; DIAGLIB_B200_FOLD_TRMM=0 DIAGLIB_B200_WS_MASK=2 python tools/determinism_check.py 128 10
 echo "== direct-store, no bulk-copy Gram (mask 4)"; DIAGLIB_B200_FOLD_TRMM=0 DIAGLIB_B200_WS_MASK=4 python tools/determinism_check.py 128 10) > gpurun_out/dbg5.log 2>&1
cp /tmp/shipped.so diaglib_b200/libdiaglib_b200.so
grep "^==\|DETERM\|rep 0" gpurun_out/dbg5.log; grep -c "ok=False" gpurun_out/dbg5.log

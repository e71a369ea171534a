#!/bin/bash
mkdir -p gpurun_out
cp diaglib_b200/libdiaglib_b200.so /tmp/shipped.so
cp build/dbg_1.so diaglib_b200/libdiaglib_b200.so
F="grep -v ^rep.[1-9].*sig=bc6bb35fe7ec"
(echo "== direct-store, two-sided single-CTA eigensolver"; DIAGLIB_B200_FOLD_TRMM=0 DIAGLIB_B200_EIG_MODE=1 python tools/determinism_check.py 128 10
 echo "== direct-store, no TMA Gram"; DIAGLIB_B200_FOLD_TRMM=0 DIAGLIB_B200_NO_TMA=1 python tools/determinism_check.py 128 10
 echo "== direct-store, no warp-specialised kernels at all"; DIAGLIB_B200_FOLD_TRMM=0 DIAGLIB_B200_NO_WS=1 python tools/determinism_check.py 128 10
 echo "== direct-store, natural row order"; DIAGLIB_B200_FOLD_TRMM=0 DIAGLIB_B200_DET_NATURAL=1 python tools/determinism_check.py 128 10) > gpurun_out/dbg4.log 2>&1
cp /tmp/shipped.so diaglib_b200/libdiaglib_b200.so
grep -v "^rep [1-9].*ok=True its=24 passes=244 sweeps=90" gpurun_out/dbg4.log

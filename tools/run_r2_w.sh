#!/bin/bash
mkdir -p gpurun_out
(echo "== fold, chain"; python tools/determinism_check.py 128 5
 echo "== no fold, chain"; DIAGLIB_B200_FOLD_TRMM=0 python tools/determinism_check.py 128 5
 echo "== fold, host-driven"; DIAGLIB_B200_SPEC_ORTHO=0 python tools/determinism_check.py 128 4
 echo "== no fold, host-driven"; DIAGLIB_B200_SPEC_ORTHO=0 DIAGLIB_B200_FOLD_TRMM=0 python tools/determinism_check.py 128 4
 echo "== no fold, no ident"; DIAGLIB_B200_NO_IDENT_PROJ=1 DIAGLIB_B200_FOLD_TRMM=0 python tools/determinism_check.py 128 4) > gpurun_out/w_det.log 2>&1
cat gpurun_out/w_det.log

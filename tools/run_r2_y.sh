#!/bin/bash
mkdir -p gpurun_out
(echo "== fold (default)"; python tools/determinism_check.py 128 10
 echo "== fold off"; DIAGLIB_B200_FOLD_TRMM=0 python tools/determinism_check.py 128 10
 echo "== fold, host-driven"; DIAGLIB_B200_SPEC_ORTHO=0 python tools/determinism_check.py 128 6) > gpurun_out/y_det.log 2>&1
cat gpurun_out/y_det.log | grep -v "^rep" ; grep -c "^rep" gpurun_out/y_det.log

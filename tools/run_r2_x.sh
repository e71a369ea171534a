#!/bin/bash
mkdir -p gpurun_out
(for w in d c a b; do echo "== worktree $w"; (cd _wt/$w && python tools/determinism_check.py 128 8); done
 echo "== current, fold off"; DIAGLIB_B200_FOLD_TRMM=0 python tools/determinism_check.py 128 8) > gpurun_out/x_det.log 2>&1
cat gpurun_out/x_det.log

#!/bin/bash
# race localisation: epilogue variants of the block multiply (build/dbg_N.so), fold off so that the separate triangular
# multiply runs as in the builds that showed the problem
mkdir -p gpurun_out
cp diaglib_b200/libdiaglib_b200.so /tmp/shipped.so
(for v in 1 2 3 4 5 6; do
  cp build/dbg_$v.so diaglib_b200/libdiaglib_b200.so
  echo "== variant $v"; DIAGLIB_B200_FOLD_TRMM=0 python tools/determinism_check.py 128 8 | grep -v "^rep [1-7].*sig=bc6bb35fe7ec"
done
cp /tmp/shipped.so diaglib_b200/libdiaglib_b200.so) > gpurun_out/dbg_det.log 2>&1
cat gpurun_out/dbg_det.log

#!/bin/bash
mkdir -p gpurun_out
(echo "== shipped (one CTA per SM), fold"; python tools/determinism_check.py 128 12
 echo "== shipped, fold off"; DIAGLIB_B200_FOLD_TRMM=0 python tools/determinism_check.py 128 8) > gpurun_out/fix_det.log 2>&1
grep "^==\|DETERM" gpurun_out/fix_det.log
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/fix_bench.json 2> gpurun_out/fix_bench.err
echo "bench rc=$?"
timeout 900 python -m pytest tests/test_gpu_drivers.py tests/test_gpu_kernels.py -m gpu -q --timeout 600 --timeout-method=thread 2>&1 | tail -3

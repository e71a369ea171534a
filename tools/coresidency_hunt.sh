#!/bin/bash
# Next steps for the open question of DESIGN.md section 3 (two co-resident CTAs of blockmul_ws_kernel make a solve
# irreproducible; one CTA per SM does not).  Each step is one short GPU job; results go to gpurun_out/hunt_*.log.
#
#   1. reproduce with the shipped epilogue at the old placement (rare) and with the direct-store epilogue (frequent):
#        DIAGLIB_B200_DBG=32 DIAGLIB_B200_BMUL_RT256=0 python tools/determinism_check.py 128 20
#        (direct-store build: nvcc ... -DDLB_DBG=1 -c dense.cu, link as build/dbg_1.so, copy over the library)
#   2. compute-sanitizer on a small instance of the same configuration (n >= 512 keeps the warp-specialised kernels):
#        compute-sanitizer --tool racecheck  python tools/determinism_check.py 32 1     (shared-memory hazards)
#        compute-sanitizer --tool synccheck  python tools/determinism_check.py 32 1     (barrier misuse)
#        compute-sanitizer --tool initcheck  python tools/determinism_check.py 32 1     (uninitialised global reads)
#   3. which call shape: tools/kernel_repro.py repeats single kernels at ONE shape each; extend it with the shapes of a
#      solve (q = n_act in 8..36, p = 37 + 2 n_act) and with back-to-back pairs (Ritz product -> residual, projection ->
#      Gram) before concluding that the kernel alone is reproducible.
#   4. per-call checksums inside a solve: set DIAGLIB_B200_DBG to a (new) bit that makes launch_blockmul follow every
#      call with a Gram-free checksum kernel of Y into a log; the first call whose checksum differs between two runs
#      names the shape and the neighbours.
#   5. hardware-side questions that remained open: does the block scheduler ever place two of these CTAs on one SM
#      when the grid is 148 (the shipped request is padded to 116 KB to exclude it); does the problem follow the SM
#      (smid logged per CTA) or the CTA pair.
set -e
mkdir -p gpurun_out
DIAGLIB_B200_DBG=32 DIAGLIB_B200_BMUL_RT256=0 python tools/determinism_check.py 128 20 > gpurun_out/hunt_1.log 2>&1 || true
for t in racecheck synccheck initcheck; do
  DIAGLIB_B200_DBG=32 DIAGLIB_B200_BMUL_RT256=0 timeout 900 compute-sanitizer --tool $t python tools/determinism_check.py 32 1 > gpurun_out/hunt_2_$t.log 2>&1 || true
done
tail -n 3 gpurun_out/hunt_*.log

#!/bin/bash
# round-2 GPU session A: new eigensolver tests + timing, full GPU suite, the real C3 oracle solve on the host, a bench line
mkdir -p gpurun_out
nproc > gpurun_out/a_nproc.txt; free -g >> gpurun_out/a_nproc.txt
(python bench.py --impl reference > gpurun_out/ref_c3.json 2> gpurun_out/ref_c3.err; echo "ref rc=$?" >> gpurun_out/ref_c3.err) &
REFPID=$!
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q -s -k "sym_eig" --timeout 300 --timeout-method=thread > gpurun_out/a_eig.log 2>&1
echo "eig rc=$?" >> gpurun_out/a_eig.log
timeout 1500 python -m pytest tests -m gpu -x -q --timeout 600 --timeout-method=thread > gpurun_out/a_all.log 2>&1
echo "all rc=$?" >> gpurun_out/a_all.log
wait $REFPID
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/a_bench.json 2> gpurun_out/a_bench.err
echo "bench rc=$?" >> gpurun_out/a_bench.err
tail -3 gpurun_out/a_eig.log gpurun_out/a_all.log gpurun_out/ref_c3.err gpurun_out/a_bench.err

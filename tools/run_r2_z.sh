#!/bin/bash
# round-2 GPU session Z: final build - full GPU suite, bench line, ncu launch list of the bench command
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 --timeout-method=thread > gpurun_out/z_gpu_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/z_gpu_tests.log
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/z_bench.json 2> gpurun_out/z_bench.err
echo "bench rc=$?" >> gpurun_out/z_bench.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4300 --csv --log-file gpurun_out/z_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/z_ncu.log 2>&1
echo "ncu rc=$?" >> gpurun_out/z_ncu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/z_smoke.log 2>&1
tail -n 3 gpurun_out/z_gpu_tests.log gpurun_out/z_bench.err gpurun_out/z_ncu.log gpurun_out/z_smoke.log

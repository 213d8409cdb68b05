#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests -m gpu -q --maxfail=20 > gpurun_out/r02_pytest8.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_pytest8.log
timeout -k 10 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r02_bench8.json 2> gpurun_out/r02_bench8.err
echo "bench rc=$?" >> gpurun_out/r02_bench8.err
timeout -k 10 400 ncu --set full --clock-control none --import-source on -k regex:radial_integral_runs_kernel -c 1 -f -o gpurun_out/r02_radial_runs python tools/prof_c2.py > gpurun_out/r02_ncu_c2.log 2>&1
tail -n 5 gpurun_out/r02_pytest8.log; tail -n 2 gpurun_out/r02_bench8.err; tail -n 3 gpurun_out/r02_ncu_c2.log

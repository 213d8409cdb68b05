#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests/test_stein_gpu.py tests/test_twostage_gpu.py tests/test_pair_gpu.py -m gpu -q --maxfail=20 > gpurun_out/r02_pytest4.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_pytest4.log
timeout -k 10 400 python tools/ab_stein.py 296 > gpurun_out/r02_ab_stein.json 2> gpurun_out/r02_ab_stein.err
echo "ab rc=$?" >> gpurun_out/r02_ab_stein.err
timeout -k 10 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_c3_launches_stein.csv python tools/prof_c3.py 148 > gpurun_out/r02_ncu_list4.log 2>&1
timeout -k 10 900 python -m pytest tests -m gpu -q --maxfail=20 > gpurun_out/r02_pytest4_full.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_pytest4_full.log
cat gpurun_out/r02_ab_stein.json; tail -n 3 gpurun_out/r02_ab_stein.err
tail -n 12 gpurun_out/r02_pytest4.log; tail -n 5 gpurun_out/r02_pytest4_full.log

#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests -m gpu -q --maxfail=20 > gpurun_out/r02_pytest10.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_pytest10.log
timeout -k 10 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r02_bench10.json 2> gpurun_out/r02_bench10.err
echo "bench rc=$?" >> gpurun_out/r02_bench10.err
timeout -k 10 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r02_c5_launches.csv python tools/dev_c5.py 32 250 > gpurun_out/r02_ncu_c5.log 2>&1
tail -n 4 gpurun_out/r02_pytest10.log; tail -n 2 gpurun_out/r02_bench10.err; tail -n 4 gpurun_out/r02_ncu_c5.log

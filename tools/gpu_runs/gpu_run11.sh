#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests -m gpu -q --maxfail=20 > gpurun_out/r02_pytest11.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_pytest11.log
timeout -k 10 600 python tools/dev_c5.py 148 250 > gpurun_out/r02_c5_148.log 2>&1
timeout -k 10 300 python tools/ab_stein.py 296 > gpurun_out/r02_ab_stein11.json 2>/dev/null
tail -n 4 gpurun_out/r02_pytest11.log; tail -n 4 gpurun_out/r02_c5_148.log; cat gpurun_out/r02_ab_stein11.json

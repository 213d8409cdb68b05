#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout -k 10 120 python tools/ab_sy2sb.py 160 2 --small > gpurun_out/r02_ab5_small.json 2> gpurun_out/r02_ab5_small.err
echo "small rc=$?" >> gpurun_out/r02_ab5_small.err
timeout -k 10 120 python tools/ab_sy2sb.py 333 5 --small > gpurun_out/r02_ab5_small2.json 2>> gpurun_out/r02_ab5_small.err
echo "small2 rc=$?" >> gpurun_out/r02_ab5_small.err
timeout -k 10 300 python tools/ab_sy2sb.py 2363 296 > gpurun_out/r02_ab5_c3.json 2> gpurun_out/r02_ab5_c3.err
echo "ab rc=$?" >> gpurun_out/r02_ab5_c3.err
timeout -k 10 600 python -m pytest tests/test_twostage_gpu.py tests/test_pair_gpu.py tests/test_stein_gpu.py -m gpu -q --maxfail=20 > gpurun_out/r02_pytest5.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_pytest5.log
cat gpurun_out/r02_ab5_small.json gpurun_out/r02_ab5_small2.json gpurun_out/r02_ab5_c3.json
tail -n 4 gpurun_out/r02_ab5_small.err gpurun_out/r02_ab5_c3.err; tail -n 8 gpurun_out/r02_pytest5.log

#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout -k 10 240 python tools/ab_sy2sb.py --small > gpurun_out/r02_ab_small.json 2> gpurun_out/r02_ab_small.err
echo "small rc=$?" >> gpurun_out/r02_ab_small.err
timeout -k 10 400 compute-sanitizer --tool racecheck --print-limit 20 python tools/ab_sy2sb.py 160 2 --small > gpurun_out/r02_racecheck.log 2>&1
echo "racecheck rc=$?" >> gpurun_out/r02_racecheck.log
timeout -k 10 400 compute-sanitizer --tool memcheck --print-limit 20 python tools/ab_sy2sb.py 160 2 --small > gpurun_out/r02_memcheck.log 2>&1
echo "memcheck rc=$?" >> gpurun_out/r02_memcheck.log
timeout -k 10 300 python tools/ab_sy2sb.py 2363 296 > gpurun_out/r02_ab_c3.json 2> gpurun_out/r02_ab_c3.err
echo "ab rc=$?" >> gpurun_out/r02_ab_c3.err
timeout -k 10 300 python tools/ab_sy2sb.py 1932 148 > gpurun_out/r02_ab_c4.json 2>> gpurun_out/r02_ab_c3.err
timeout -k 10 900 python -m pytest tests -m gpu -q --maxfail=20 > gpurun_out/r02_pytest3.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_pytest3.log
cat gpurun_out/r02_ab_small.json gpurun_out/r02_ab_c3.json gpurun_out/r02_ab_c4.json
tail -n 3 gpurun_out/r02_ab_small.err gpurun_out/r02_racecheck.log gpurun_out/r02_memcheck.log gpurun_out/r02_ab_c3.err
tail -n 6 gpurun_out/r02_pytest3.log

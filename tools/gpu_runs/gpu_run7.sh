#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout -k 10 120 python tools/ab_sy2sb.py 333 5 --small > gpurun_out/r02_ab7_small.json 2> gpurun_out/r02_ab7.err
echo "small rc=$?" >> gpurun_out/r02_ab7.err
timeout -k 10 300 python tools/ab_sy2sb.py 2363 296 > gpurun_out/r02_ab7_c3.json 2>> gpurun_out/r02_ab7.err
echo "ab rc=$?" >> gpurun_out/r02_ab7.err
cat gpurun_out/r02_ab7_small.json gpurun_out/r02_ab7_c3.json; tail -n 3 gpurun_out/r02_ab7.err

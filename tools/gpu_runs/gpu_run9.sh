#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout -k 10 600 python tools/ab_small_n.py > gpurun_out/r02_ab_small_n.json 2> gpurun_out/r02_ab_small_n.err
echo "rc=$?" >> gpurun_out/r02_ab_small_n.err
cat gpurun_out/r02_ab_small_n.json; tail -n 3 gpurun_out/r02_ab_small_n.err

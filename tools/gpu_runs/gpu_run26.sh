#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout -k 10 200 python tools/ab_multi.py 2363 296 > gpurun_out/r02_ab26.json 2>> gpurun_out/r02_ab26.err
timeout -k 10 200 python tools/ab_multi.py 1932 296 >> gpurun_out/r02_ab26.json 2>> gpurun_out/r02_ab26.err
timeout -k 10 200 python tools/ab_multi.py 1100 296 >> gpurun_out/r02_ab26.json 2>> gpurun_out/r02_ab26.err
cat gpurun_out/r02_ab26.json
timeout -k 10 600 python tools/ab_crossover.py 74 > gpurun_out/r02_crossover74.json 2>> gpurun_out/r02_ab26.err
cat gpurun_out/r02_crossover74.json
timeout -k 10 600 python tools/ab_crossover.py 592 > gpurun_out/r02_crossover592.json 2>> gpurun_out/r02_ab26.err
cat gpurun_out/r02_crossover592.json
tail -n 5 gpurun_out/r02_ab26.err

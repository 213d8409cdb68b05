#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout -k 10 400 python tools/ab_pipeline.py 125 74 250 > gpurun_out/r02_ab_pipeline.json 2> gpurun_out/r02_ab43.err; cat gpurun_out/r02_ab_pipeline.json; tail -n 3 gpurun_out/r02_ab43.err
for pl in 0 1; do
echo "pipeline $pl"; ARCB200_PIPELINE=$pl timeout -k 10 400 python tools/dev_c5.py 74 2>&1 | grep "diagonalise\|max" | tail -n 2
done
timeout -k 10 900 python -m pytest tests -m gpu -q --maxfail=20 -k "pair or round2 or c6 or stein" > gpurun_out/r02_pytest43.log 2>&1
tail -n 3 gpurun_out/r02_pytest43.log

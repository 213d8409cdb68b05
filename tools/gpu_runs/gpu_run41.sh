#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests -m gpu -q --maxfail=20 > gpurun_out/r02_pytest41.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_pytest41.log
tail -n 4 gpurun_out/r02_pytest41.log
for w in 4 2; do
echo "q2 warps $w"; ARCB200_Q2_WARPS=$w timeout -k 10 400 python tools/dev_c5.py 74 2>&1 | grep "diagonalise\|max" | tail -n 2
done

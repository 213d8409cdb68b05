#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
ARCB200_SY2SB_STREAMS=1 timeout -k 10 600 ncu --set full --clock-control none --import-source on -k regex:q2_kernel -c 1 -f -o gpurun_out/r02_q2 python tools/prof_c3.py 148 > gpurun_out/r02_ncu35.log 2>&1
tail -n 2 gpurun_out/r02_ncu35.log

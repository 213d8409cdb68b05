#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
ARCB200_SY2SB_VARIANT=ring timeout -k 10 600 ncu --set full --clock-control none --import-source on -k regex:sy2sb_ring_kernel -c 1 -f -o gpurun_out/r02_sy2sb_ring python tools/ab_multi.py 2363 148 > gpurun_out/r02_ncu_ring.log 2>&1
tail -n 3 gpurun_out/r02_ncu_ring.log

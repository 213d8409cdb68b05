#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout -k 10 600 ncu --set full --clock-control none --import-source on -k regex:sy2sb_fused_kernel -c 1 -f -o gpurun_out/r02_sy2sb_fused python tools/ab_sy2sb.py 2363 148 > gpurun_out/r02_ncu_fused.log 2>&1
tail -n 3 gpurun_out/r02_ncu_fused.log

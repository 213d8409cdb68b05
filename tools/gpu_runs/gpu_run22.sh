#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout -k 10 900 ncu --set full --clock-control none --import-source on -k regex:sb2st_split_kernel -c 1 -f -o gpurun_out/r02_sb2st_split python tools/ab_sb2st.py 2363 148 > gpurun_out/r02_ncu_sb2st.log 2>&1
tail -n 3 gpurun_out/r02_ncu_sb2st.log
ls -la gpurun_out/*.ncu-rep

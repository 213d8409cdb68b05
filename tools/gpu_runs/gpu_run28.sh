#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout -k 10 600 ncu --set full --clock-control none --import-source on -k regex:sy2sb_phase_kernel --launch-skip 32 --launch-count 1 -f -o gpurun_out/r02_syr2k_p10 python tools/prof_sy2sb_once.py 2363 148 > gpurun_out/r02_ncu28.log 2>&1
tail -n 2 gpurun_out/r02_ncu28.log
timeout -k 10 600 ncu --set full --clock-control none --import-source on -k regex:sy2sb_phase_kernel --launch-skip 30 --launch-count 1 -f -o gpurun_out/r02_symm_p10 python tools/prof_sy2sb_once.py 2363 148 >> gpurun_out/r02_ncu28.log 2>&1
tail -n 2 gpurun_out/r02_ncu28.log
ls -la gpurun_out/*.ncu-rep

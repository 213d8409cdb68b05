#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
for cfg in c1 c4; do for ts in 0 1; do
ARCB200_TWO_STAGE=$ts timeout -k 10 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/r02_small_${cfg}_ts${ts}.csv python tools/prof_small.py $cfg > gpurun_out/r02_small_${cfg}_ts${ts}.log 2>&1
echo "== $cfg two_stage=$ts"; python tools/summarize_launches.py gpurun_out/r02_small_${cfg}_ts${ts}.csv | head -14
done; done
timeout -k 10 300 python tools/ab_small_n.py

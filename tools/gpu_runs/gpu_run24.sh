#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
: > gpurun_out/r02_ab24.json
for cfg in "410 300" "651 286"; do
  for pr in 8 4 2; do
    echo "pairs=$pr" >> gpurun_out/r02_ab24.json
    ARCB200_TWO_STAGE=1 ARCB200_SB2ST_PAIRS=$pr timeout -k 10 200 python tools/ab_sb2st.py $cfg >> gpurun_out/r02_ab24.json 2>> gpurun_out/r02_ab24.err
  done
done
for q in 0 1; do
  echo "qr2=$q" >> gpurun_out/r02_ab24.json
  ARCB200_SY2SB_QR2=$q ARCB200_TWO_STAGE=1 timeout -k 10 200 python tools/ab_multi.py 410 300 >> gpurun_out/r02_ab24.json 2>> gpurun_out/r02_ab24.err
  ARCB200_SY2SB_QR2=$q ARCB200_TWO_STAGE=1 timeout -k 10 200 python tools/ab_multi.py 651 286 >> gpurun_out/r02_ab24.json 2>> gpurun_out/r02_ab24.err
  ARCB200_SY2SB_QR2=$q timeout -k 10 200 python tools/ab_multi.py 2363 250 >> gpurun_out/r02_ab24.json 2>> gpurun_out/r02_ab24.err
done
cat gpurun_out/r02_ab24.json; tail -n 5 gpurun_out/r02_ab24.err
timeout -k 10 300 python tools/ab_small_n.py
timeout -k 10 600 python tools/ab_crossover.py 296 > gpurun_out/r02_crossover.json 2>> gpurun_out/r02_ab24.err
cat gpurun_out/r02_crossover.json
timeout -k 10 900 python -m pytest tests -m gpu -q --maxfail=20 > gpurun_out/r02_pytest24.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_pytest24.log
tail -n 4 gpurun_out/r02_pytest24.log

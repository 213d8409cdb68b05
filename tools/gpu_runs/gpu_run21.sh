#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
: > gpurun_out/r02_ab21_sb2st.json
for cfg in "1027 3" "1100 20" "2363 296" "1932 148" "4096 20"; do
  timeout -k 10 200 python tools/ab_sb2st.py $cfg >> gpurun_out/r02_ab21_sb2st.json 2>> gpurun_out/r02_ab21.err
  echo "sb2st $cfg rc=$?" >> gpurun_out/r02_ab21.err
done
cat gpurun_out/r02_ab21_sb2st.json
tail -n 8 gpurun_out/r02_ab21.err
timeout -k 10 900 python -m pytest tests -m gpu -q --maxfail=20 > gpurun_out/r02_pytest21.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_pytest21.log
tail -n 4 gpurun_out/r02_pytest21.log

#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
: > gpurun_out/r02_ab19_multi.json
for cfg in "2363 125" "2363 250" "1932 100" "4096 40" "10384 8"; do
  timeout -k 10 300 python tools/ab_multi.py $cfg >> gpurun_out/r02_ab19_multi.json 2>> gpurun_out/r02_ab19.err
  echo "multi $cfg rc=$?" >> gpurun_out/r02_ab19.err
done
timeout -k 10 300 python tools/ab_sy2sb.py 2363 296 > gpurun_out/r02_ab19_c3.json 2>> gpurun_out/r02_ab19.err
echo "ab rc=$?" >> gpurun_out/r02_ab19.err
timeout -k 10 900 python -m pytest tests -m gpu -q --maxfail=20 > gpurun_out/r02_pytest19.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_pytest19.log
timeout -k 10 600 python tools/dev_c5.py > gpurun_out/r02_c5_19.log 2>&1
echo "c5 rc=$?" >> gpurun_out/r02_c5_19.log
cat gpurun_out/r02_ab19_multi.json
python - <<PY
import json
d=json.load(open("gpurun_out/r02_ab19_c3.json")); print({k:(round(v["ms"],1) if isinstance(v,dict) else v) for k,v in d.items()})
PY
tail -n 8 gpurun_out/r02_ab19.err; tail -n 4 gpurun_out/r02_pytest19.log; tail -n 6 gpurun_out/r02_c5_19.log

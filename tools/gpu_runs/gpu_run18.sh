#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout -k 10 300 python tools/ab_sy2sb.py 2363 296 > gpurun_out/r02_ab18_c3.json 2> gpurun_out/r02_ab18.err
echo "ab rc=$?" >> gpurun_out/r02_ab18.err
timeout -k 10 200 python tools/ab_sy2sb.py 1932 148 > gpurun_out/r02_ab18_c4.json 2>> gpurun_out/r02_ab18.err
timeout -k 10 900 python -m pytest tests -m gpu -q --maxfail=20 > gpurun_out/r02_pytest18.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_pytest18.log
python - <<PY
import json
for f in ("gpurun_out/r02_ab18_c3.json","gpurun_out/r02_ab18_c4.json"):
    d=json.load(open(f)); print({k:(round(v["ms"],1) if isinstance(v,dict) else v) for k,v in d.items() if k in ("N","batch","lockstep","wg","fused","ring","eig_err_worst_of_3_matrices_wg_fused")})
PY
tail -n 2 gpurun_out/r02_ab18.err; tail -n 4 gpurun_out/r02_pytest18.log

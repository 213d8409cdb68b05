#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout -k 10 300 python tools/ab_sy2sb.py 2363 296 > gpurun_out/r02_ab36_c3.json 2>> gpurun_out/r02_ab36.err
python - <<PY
import json
d=json.load(open("gpurun_out/r02_ab36_c3.json")); print({k:(round(v["ms"],1) if isinstance(v,dict) else v) for k,v in d.items()})
PY
AB_ONLY=multi timeout -k 10 200 python tools/ab_multi.py 2363 125; AB_ONLY=multi timeout -k 10 200 python tools/ab_multi.py 651 286
timeout -k 10 900 python -m pytest tests -m gpu -q --maxfail=20 -k "twostage or pair or stark or round2" > gpurun_out/r02_pytest36.log 2>&1
tail -n 3 gpurun_out/r02_pytest36.log

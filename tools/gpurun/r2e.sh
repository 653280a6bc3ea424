#!/bin/bash
# round 2, call E: new parity tests (planted retrieval ranks, evaluate_vl vs oracle, medium/large, GELU module)
cd /root/repo
timeout 1200 python -m pytest tests/test_engine_gpu.py tests/test_model_gpu.py -q -x -k "planted or evaluate_vl or deeper or gelu" > gpurun_out/r2e_tests.log 2>&1; echo "tests rc=$?"; tail -25 gpurun_out/r2e_tests.log
python - <<'PY'
import json
try:
    rows = json.load(open("gpurun_out/planted_retrieval.json"))
    print([(r["planted"], r["rank_gpu"], r["rank_oracle"], round(r["max_err"], 3), round(r["margin"], 2)) for r in rows])
except Exception as e:
    print("ERR", e)
PY

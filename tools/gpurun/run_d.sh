#!/bin/bash
# call D: LN-bwd knob sweep + tests + bench
cd /root/repo
python tools/ln_sweep.py > gpurun_out/d_ln.log 2>&1
for p in 2 8 16 32; do MVLT_LN_BWD_PASSES=$p python tools/ln_sweep.py >> gpurun_out/d_ln.log 2>&1; done
MVLT_LN_BWD_VEC=0 python tools/ln_sweep.py >> gpurun_out/d_ln.log 2>&1
MVLT_LN_BWD_BPS=8 python tools/ln_sweep.py >> gpurun_out/d_ln.log 2>&1
MVLT_LN_BWD_BPS=2 MVLT_LN_BWD_PASSES=8 python tools/ln_sweep.py >> gpurun_out/d_ln.log 2>&1
cat gpurun_out/d_ln.log
timeout 300 python -m pytest tests -m gpu -q -x > gpurun_out/d_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/d_tests.log
timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu --retrieval-queries 0 > gpurun_out/d_bench.json 2>gpurun_out/d_bench.err
python - <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/d_bench.json").read().strip().splitlines()[-1])
    print("bench", d["value"], d["ms_per_step"], d["e2e"]["value"], d["host_enqueue_ms_per_step"])
    print({k: v["ms_per_step"] for k, v in d["kernel_breakdown"].items()})
except Exception as e:
    print("ERR", e); print(open("gpurun_out/d_bench.err").read()[-1500:])
PY

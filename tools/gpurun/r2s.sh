#!/bin/bash
# round 2, call S: C = 128 fused MLP backward: unit tests, timing, model parity, bench
cd /root/repo
timeout 200 python -m pytest tests/test_mlp_gpu.py -q -x 2>&1 | tail -8
rc=${PIPESTATUS[0]}
timeout 200 python tools/mlp_bench.py --bwd 2>&1 | tail -4
timeout 900 python -m pytest tests/test_model_gpu.py -q -x 2>&1 | tail -3
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --no-eager --retrieval-queries 0 > gpurun_out/r2s_bench.json 2> gpurun_out/r2s_bench.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2s_bench.json").read().strip().splitlines()[-1])
print("train", d["value"], d["ms_per_step"], d["e2e"]["value"], d["host_enqueue_ms_per_step"])
print({k: (v["ms_per_step"], v["launches_per_step"]) for k, v in list(d["kernel_breakdown"].items())[:16]})
print("sub", {k: (v.get("value"), v.get("e2e", {}).get("value")) for k, v in (d.get("sub_benches") or {}).items()})
PY

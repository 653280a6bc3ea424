#!/bin/bash
# round 2, call I: fused MLP backward: unit tests under a timeout, model parity, timing, bench
cd /root/repo
timeout 200 python -m pytest tests/test_mlp_gpu.py -q -x 2>&1 | tail -12
rc=${PIPESTATUS[0]}
if [ "$rc" = "0" ]; then
  timeout 200 python tools/mlp_bench.py --bwd 2>&1 | tail -4
  timeout 900 python -m pytest tests/test_model_gpu.py tests/test_engine_gpu.py -q -x 2>&1 | tail -6
  timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --no-eager --no-sub --retrieval-queries 0 > gpurun_out/r2i_bench.json 2> gpurun_out/r2i_bench.err
  python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2i_bench.json").read().strip().splitlines()[-1])
print("train", d["value"], d["ms_per_step"], d["e2e"]["value"], d["host_enqueue_ms_per_step"])
print({k: (v["ms_per_step"], v["launches_per_step"]) for k, v in d["kernel_breakdown"].items()})
PY
fi

#!/bin/bash
# round 2, call H: fused MLP forward: tests, timing vs the two-GEMM path, model tests (inference path uses it), retrieval bench
cd /root/repo
timeout 300 python -m pytest tests/test_mlp_gpu.py -q -x 2>&1 | tail -3
timeout 300 python tools/mlp_bench.py 2>&1 | tail -4
timeout 600 python -m pytest tests/test_model_gpu.py tests/test_engine_gpu.py -q -x -k "forward_logits or retrieval or recognition or evaluate_vl" 2>&1 | tail -3
timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu --no-eager --no-sub > gpurun_out/r2h_bench.json 2> gpurun_out/r2h_bench.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2h_bench.json").read().strip().splitlines()[-1])
r = d["retrieval"]
print("train", d["value"], "retrieval", r["value"], r.get("e2e"), r.get("roofline", {}).get("gemm_share_of_step"))
print({k: (v["ms_per_step"], v["launches_per_step"]) for k, v in r.get("kernel_breakdown", {}).items()})
PY

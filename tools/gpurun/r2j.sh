#!/bin/bash
# round 2, call J: 16 GELU warps in the fused MLP kernels: tests, timing, bench
cd /root/repo
timeout 200 python -m pytest tests/test_mlp_gpu.py -q -x 2>&1 | tail -3
timeout 200 python tools/mlp_bench.py --bwd 2>&1 | tail -4
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --no-eager --no-sub --retrieval-queries 200 > gpurun_out/r2j_bench.json 2> gpurun_out/r2j_bench.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2j_bench.json").read().strip().splitlines()[-1])
print("train", d["value"], d["ms_per_step"], d["e2e"]["value"], d["host_enqueue_ms_per_step"], "retr", d["retrieval"]["value"])
print({k: (v["ms_per_step"], v["launches_per_step"]) for k, v in d["kernel_breakdown"].items()})
PY

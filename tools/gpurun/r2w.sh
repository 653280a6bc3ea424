#!/bin/bash
# round 2, call W: 4x-unrolled layout kernels, 16-byte BatchNorm reductions: kernel tests, model tests, bench
cd /root/repo
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_gemm_gpu.py -q -x 2>&1 | tail -3
timeout 900 python -m pytest tests/test_model_gpu.py -q -x -k "not b128 and not deeper" 2>&1 | tail -3
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --no-eager --no-sub --retrieval-queries 200 > gpurun_out/r2w_bench.json 2> gpurun_out/r2w_bench.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2w_bench.json").read().strip().splitlines()[-1])
print("train", d["value"], d["ms_per_step"], d["e2e"]["value"], d["host_enqueue_ms_per_step"], "retr", d["retrieval"]["value"])
print({k: (v["ms_per_step"], v["launches_per_step"]) for k, v in list(d["kernel_breakdown"].items())[:30]})
PY

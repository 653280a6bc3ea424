#!/bin/bash
# round 2, call V: 8-warp attention backward, LayerNorm retune, head selection: tests, smoke, bench
cd /root/repo
timeout 200 python -m pytest tests/test_attention_gpu.py -q -x 2>&1 | tail -3
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r2v_tests.log 2>&1; echo "gpu tests rc=$?"; tail -3 gpurun_out/r2v_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --no-eager --no-sub --retrieval-queries 200 > gpurun_out/r2v_bench.json 2> gpurun_out/r2v_bench.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2v_bench.json").read().strip().splitlines()[-1])
print("train", d["value"], d["ms_per_step"], d["e2e"]["value"], d["host_enqueue_ms_per_step"], "retr", d["retrieval"]["value"])
print({k: (v["ms_per_step"], v["launches_per_step"]) for k, v in list(d["kernel_breakdown"].items())[:12]})
PY

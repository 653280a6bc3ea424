#!/bin/bash
# round 2 (session 2), call P: first block's norm1 chained onto the stage-embedding LayerNorms: tests, bench
cd /root/repo
timeout 300 python -m pytest tests/test_kernels_gpu.py -q -x -k "layernorm" > gpurun_out/r3p_t1.log 2>&1; echo "ln tests rc=$?"; tail -5 gpurun_out/r3p_t1.log
timeout 900 python -m pytest tests/test_model_gpu.py tests/test_graph_gpu.py tests/test_engine_gpu.py -q -x > gpurun_out/r3p_t2.log 2>&1; echo "model tests rc=$?"; tail -4 gpurun_out/r3p_t2.log
timeout 400 python bench.py --steps 20 --warmup 5 --no-cpu --no-eager --no-sub > gpurun_out/r3p_bench.json 2> gpurun_out/r3p_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/r3p_bench.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r3p_bench.json").read().strip().splitlines()[-1])
r = d["retrieval"]
print(d["value"], d["ms_per_step"], d["e2e"]["value"], "retr", r["value"], r["e2e"]["value"], "lnfwd", d["kernel_breakdown"].get("layernorm_fwd"))
print({k: v["ms_per_step"] for k, v in r["kernel_breakdown"].items()})
PY

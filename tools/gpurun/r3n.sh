#!/bin/bash
# round 2 (session 2), call N: LayerNorm fused into the residual GEMM epilogue (C = 64 / 128): kernel test, model tests, A/B in the bench
cd /root/repo
timeout 300 python -m pytest tests/test_gemm_gpu.py -q -x -k "layernorm or epilogues" > gpurun_out/r3n_t1.log 2>&1; echo "gemm ln tests rc=$?"; tail -6 gpurun_out/r3n_t1.log
timeout 900 python -m pytest tests/test_model_gpu.py tests/test_graph_gpu.py -q -x > gpurun_out/r3n_t2.log 2>&1; echo "model tests rc=$?"; tail -4 gpurun_out/r3n_t2.log
for v in 1 0; do
MVLT_FUSED_LN=$v timeout 400 python bench.py --steps 20 --warmup 5 --no-cpu --no-eager --no-sub > gpurun_out/r3n_bench_ln$v.json 2> gpurun_out/r3n_bench_ln$v.err; echo "bench ln=$v rc=$?"; tail -2 gpurun_out/r3n_bench_ln$v.err
done
python - <<'PY'
import json
for n in ("ln1", "ln0"):
    try:
        d = json.loads(open(f"gpurun_out/r3n_bench_{n}.json").read().strip().splitlines()[-1])
        r = d["retrieval"]
        print(n, d["value"], d["ms_per_step"], d["e2e"]["value"], "retr", r["value"], r["e2e"]["value"], "gemm", d["roofline"]["gemm_ms_per_step"], d["roofline"]["frac_of_own_roofline"], "lnfwd", d["kernel_breakdown"].get("layernorm_fwd"))
    except Exception as e:
        print(n, "ERR", e)
PY

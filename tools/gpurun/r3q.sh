#!/bin/bash
# round 2 (session 2), call Q: spatial-reduction convolution through the 5-D TMA patch view: kernel test, model tests, A/B in the bench
cd /root/repo
timeout 300 python -m pytest tests/test_gemm_gpu.py -q -x -k "patch_view or layernorm" > gpurun_out/r3q_t1.log 2>&1; echo "patch tests rc=$?"; tail -8 gpurun_out/r3q_t1.log
timeout 900 python -m pytest tests/test_model_gpu.py tests/test_graph_gpu.py tests/test_engine_gpu.py -q -x > gpurun_out/r3q_t2.log 2>&1; echo "model tests rc=$?"; tail -4 gpurun_out/r3q_t2.log
for v in 1 0; do
MVLT_PATCH_VIEW=$v timeout 400 python bench.py --steps 20 --warmup 5 --no-cpu --no-eager --no-sub > gpurun_out/r3q_bench_pv$v.json 2> gpurun_out/r3q_bench_pv$v.err; echo "bench pv=$v rc=$?"; tail -2 gpurun_out/r3q_bench_pv$v.err
done
python - <<'PY'
import json
for n in ("pv1", "pv0"):
    try:
        d = json.loads(open(f"gpurun_out/r3q_bench_{n}.json").read().strip().splitlines()[-1])
        r = d["retrieval"]
        print(n, d["value"], d["ms_per_step"], d["e2e"]["value"], "retr", r["value"], r["e2e"]["value"], {k: v["ms_per_step"] for k, v in r["kernel_breakdown"].items() if k in ("gemm", "patchify")})
    except Exception as e:
        print(n, "ERR", e)
PY

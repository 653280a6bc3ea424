#!/bin/bash
# round 2 (session 2), call R: input gradient of the spatial-reduction convolution stored through the 5-D view: tests, A/B
cd /root/repo
timeout 300 python -m pytest tests/test_gemm_gpu.py -q -x -k "patch" > gpurun_out/r3r_t1.log 2>&1; echo "patch tests rc=$?"; tail -8 gpurun_out/r3r_t1.log
timeout 900 python -m pytest tests/test_model_gpu.py tests/test_graph_gpu.py -q -x > gpurun_out/r3r_t2.log 2>&1; echo "model tests rc=$?"; tail -4 gpurun_out/r3r_t2.log
for v in 1 0; do
MVLT_PATCH_STORE=$v timeout 400 python bench.py --steps 20 --warmup 5 --no-cpu --no-eager --no-sub --retrieval-queries 0 > gpurun_out/r3r_bench_ps$v.json 2> gpurun_out/r3r_bench_ps$v.err; echo "bench ps=$v rc=$?"; tail -2 gpurun_out/r3r_bench_ps$v.err
done
python - <<'PY'
import json
for n in ("ps1", "ps0"):
    try:
        d = json.loads(open(f"gpurun_out/r3r_bench_{n}.json").read().strip().splitlines()[-1])
        print(n, d["value"], d["ms_per_step"], d["e2e"]["value"], {k: v["ms_per_step"] for k, v in d["kernel_breakdown"].items() if k in ("gemm", "unpatchify", "patchify")})
    except Exception as e:
        print(n, "ERR", e)
PY

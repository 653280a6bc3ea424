#!/bin/bash
# round 2, call A: first device run of the fused attention backward (csrc/attn_bwd_tcgen05.cu), then A/B in the bench
cd /root/repo
MVLT_FUSED_ATTN_BWD=1 timeout 150 python -m pytest tests/test_attention_gpu.py -q -x -k backward > gpurun_out/r2a_attn_bwd.log 2>&1
rc=$?; echo "fused attention backward kernel tests rc=$rc"; tail -15 gpurun_out/r2a_attn_bwd.log
if [ $rc -eq 0 ]; then
  MVLT_FUSED_ATTN_BWD=1 timeout 300 python -m pytest tests/test_model_gpu.py tests/test_engine_gpu.py -q -x > gpurun_out/r2a_model_bwd.log 2>&1
  echo "model tests with the fused backward rc=$?"; tail -4 gpurun_out/r2a_model_bwd.log
  MVLT_FUSED_ATTN_BWD=1 timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu --retrieval-queries 0 > gpurun_out/r2a_bench_bwd1.json 2> gpurun_out/r2a_bench_bwd1.err
  timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu --retrieval-queries 0 > gpurun_out/r2a_bench_bwd0.json 2> gpurun_out/r2a_bench_bwd0.err
  python - <<'PY'
import json
for n in ("bwd0", "bwd1"):
    try:
        d = json.loads(open(f"gpurun_out/r2a_bench_{n}.json").read().strip().splitlines()[-1])
        print(n, d["value"], d["ms_per_step"], d["e2e"]["value"], d.get("hbm_bound_kernels"))
    except Exception as e:
        print(n, "ERR", e)
PY
fi
nvidia-smi --query-gpu=name,memory.used --format=csv

#!/bin/bash
# round 2, call M: (1) relu-form GELU + division-free MLP loops with the pair mode OFF, (2) CTA-pair GEMM (cta_group::2) ON
cd /root/repo
echo "== pair off"
MVLT_GEMM_PAIR=0 timeout 300 python -m pytest tests/test_mlp_gpu.py tests/test_gemm_gpu.py -q -x 2>&1 | tail -3
MVLT_GEMM_PAIR=0 timeout 200 python tools/mlp_bench.py --bwd 2>&1 | tail -4
for cfg in "24576 2048 512 gelu" "24576 512 2048 res" "49152 1280 320 gelu" "49152 320 1280 res" "16384 30528 768 plain"; do
  set -- $cfg
  for pr in 0 1; do
    echo "-- gemm $cfg pair=$pr"; MVLT_GEMM_PAIR=$pr timeout 60 python tools/gemm_time.py $1 $2 $3 $4 2>&1 | tail -2
  done
done
echo "== pair on"
timeout 300 python -m pytest tests/test_gemm_gpu.py -q -x 2>&1 | tail -5
timeout 900 python -m pytest tests/test_model_gpu.py -q -x 2>&1 | tail -3
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --no-eager --no-sub --retrieval-queries 200 > gpurun_out/r2m_bench.json 2> gpurun_out/r2m_bench.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2m_bench.json").read().strip().splitlines()[-1])
print("train", d["value"], d["ms_per_step"], d["e2e"]["value"], d["host_enqueue_ms_per_step"], "retr", d["retrieval"]["value"])
print({k: (v["ms_per_step"], v["launches_per_step"]) for k, v in d["kernel_breakdown"].items()})
print(d["roofline"]["tensor_bound_launches"], d["roofline"]["frac"], d["roofline"]["frac_of_own_roofline"])
PY

#!/bin/bash
# round 2, call T: merged N=128 H/dH MMAs in the C=64 MLP backward; A/B of the stage-2 fused MLP in training (same box)
cd /root/repo
timeout 200 python -m pytest tests/test_mlp_gpu.py -q -x 2>&1 | tail -3
timeout 200 python tools/mlp_bench.py --bwd 2>&1 | tail -4
for dims in "64,128" "64"; do
  MVLT_FUSED_MLP_TRAIN=$dims timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --no-eager --no-sub --retrieval-queries 0 > gpurun_out/r2t_bench.json 2> gpurun_out/r2t_bench.err
  python - "$dims" <<'PY'
import json, sys
d = json.loads(open("gpurun_out/r2t_bench.json").read().strip().splitlines()[-1])
kb = d["kernel_breakdown"]
print("train dims", sys.argv[1], d["value"], d["ms_per_step"], "gemm", kb["gemm"]["ms_per_step"], "mlp", kb["mlp_fwd"]["ms_per_step"], kb["mlp_bwd"]["ms_per_step"])
PY
done

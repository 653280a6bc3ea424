#!/bin/bash
# round 2, call N: ncu --set full of the fused MLP backward and of the stage-1 LayerNorm kernels; step bench with the pair mode off
cd /root/repo
timeout 300 ncu --set full --clock-control none --import-source on -k regex:mlp_bwd -s 1 -c 1 -f -o gpurun_out/r2n_mlp_bwd python tools/mlp_bench.py --bwd --iters 1 > gpurun_out/r2n_ncu1.log 2>&1; echo "ncu mlp_bwd rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:mlp_fwd -s 2 -c 1 -f -o gpurun_out/r2n_mlp_fwd python tools/mlp_bench.py --iters 1 > gpurun_out/r2n_ncu2.log 2>&1; echo "ncu mlp_fwd rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:ln_bwd -s 6 -c 2 -f -o gpurun_out/r2n_ln_bwd python tools/profile_step.py > gpurun_out/r2n_ncu3.log 2>&1; echo "ncu ln_bwd rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:ln_fwd -s 8 -c 2 -f -o gpurun_out/r2n_ln_fwd python tools/profile_step.py > gpurun_out/r2n_ncu4.log 2>&1; echo "ncu ln_fwd rc=$?"
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --no-eager --no-sub --retrieval-queries 0 > gpurun_out/r2n_bench.json 2> gpurun_out/r2n_bench.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2n_bench.json").read().strip().splitlines()[-1])
print("train", d["value"], d["ms_per_step"], d["e2e"]["value"], d["host_enqueue_ms_per_step"])
print({k: (v["ms_per_step"], v["launches_per_step"]) for k, v in d["kernel_breakdown"].items()})
PY

#!/bin/bash
# round 2, call O: A/B in ONE run: current library vs the same sources with the pre-pair gemm_tcgen05.cu (is the refactor slower?)
cd /root/repo
for rep in 1 2; do
for lib in "" "/root/repo/mvlt_b200/lib/libmvlt_b200_oldgemm.so"; do
  MVLT_LIB=$lib timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --no-eager --no-sub --retrieval-queries 0 > gpurun_out/r2o_bench.json 2> gpurun_out/r2o_bench.err
  python - "$lib" <<'PY'
import json, sys
d = json.loads(open("gpurun_out/r2o_bench.json").read().strip().splitlines()[-1])
kb = d["kernel_breakdown"]
print(sys.argv[1] or "current", "train", d["value"], d["ms_per_step"], "gemm", kb["gemm"]["ms_per_step"], "own_roofline", d["roofline"]["frac_of_own_roofline"])
PY
done
done
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2o_bench.json").read().strip().splitlines()[-1])
print({k: (v["ms_per_step"], v["launches_per_step"]) for k, v in d["kernel_breakdown"].items()})
PY

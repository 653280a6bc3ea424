#!/bin/bash
# round 2 (session 2), call L: default bench with the instrumented pass on one stream behind a spin kernel (event intervals = kernel durations)
cd /root/repo
( time timeout 900 python bench.py --no-cpu --no-eager > gpurun_out/r3l_bench.json 2> gpurun_out/r3l_bench.err ) 2>&1 | tail -3; tail -3 gpurun_out/r3l_bench.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r3l_bench.json").read().strip().splitlines()[-1])
print("train", d["value"], d["ms_per_step"], d["e2e"], d["host_enqueue_ms_per_step"], d["gpu_launches"], d["config"].get("cuda_graph"))
print("roofline", {k: d["roofline"][k] for k in ("achieved", "frac", "traffic", "frac_of_own_roofline", "tensor_bound_launches", "hbm_bound_launches", "gemm_ms_per_step", "gemm_share_of_step")})
print({k: (v["ms_per_step"], v["launches_per_step"]) for k, v in list(d["kernel_breakdown"].items())[:16]})
r = d["retrieval"]; print("retr", r["value"], r.get("e2e", {}).get("value"), {k: r["roofline"][k] for k in ("frac", "frac_of_own_roofline", "gemm_ms_per_step")})
print("hbm", {k: v["frac_of_hbm_peak"] for k, v in d.get("hbm_bound_kernels", {}).items()})
print("retr hbm", {k: v["frac_of_hbm_peak"] for k, v in r.get("hbm_bound_kernels", {}).items()})
PY

#!/bin/bash
# round 2 (session 2), call S: FINAL state (LayerNorm fusion, patch view): full GPU suite, smoke, default bench (all legs), reference arm
cd /root/repo
( time timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/r3s_tests.log 2>&1 ) 2>&1 | tail -3; echo "gpu tests rc=$?"; tail -3 gpurun_out/r3s_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
( time timeout 900 python bench.py > gpurun_out/r3s_bench.json 2> gpurun_out/r3s_bench.err ) 2>&1 | tail -3; tail -3 gpurun_out/r3s_bench.err
( time timeout 600 python bench.py --impl reference > gpurun_out/r3s_bench_ref.json 2>/dev/null ) 2>&1 | tail -3
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r3s_bench.json").read().strip().splitlines()[-1])
print("train", d["value"], d["ms_per_step"], d["e2e"], d["host_enqueue_ms_per_step"], d["gpu_launches"], d["clocks"], d["config"].get("cuda_graph"))
print("roofline", {k: d["roofline"][k] for k in ("achieved", "frac", "traffic", "frac_of_own_roofline", "tensor_bound_launches", "hbm_bound_launches", "gemm_ms_per_step")})
r = d["retrieval"]; print("retr", r["value"], r.get("e2e", {}).get("value"), r.get("cpu_baseline"))
print("sub", {k: (v.get("value"), v.get("ms_per_step"), v.get("e2e", {}).get("value")) for k, v in (d.get("sub_benches") or {}).items()})
print("eager", d.get("gpu_eager_reference", {}).get("value"), d.get("gpu_eager_reference", {}).get("ours_over_eager")); print("cpu", d.get("cpu_baseline", {}).get("value"))
print("hbm", {k: v["frac_of_hbm_peak"] for k, v in d.get("hbm_bound_kernels", {}).items()})
r = json.loads(open("gpurun_out/r3s_bench_ref.json").read().strip().splitlines()[-1]); print("ref arm", r["value"], r["cpu_baseline"]["kind"], r["cpu_baseline"]["cores"])
PY

#!/bin/bash
# round 2 (session 2), call C: full GPU suite with the graph-mode train loop, smoke, default bench
cd /root/repo
( time timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r3c_tests.log 2>&1 ) 2>&1 | tail -3; echo "gpu tests rc=$?"; tail -8 gpurun_out/r3c_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
( time timeout 900 python bench.py > gpurun_out/r3c_bench.json 2> gpurun_out/r3c_bench.err ) 2>&1 | tail -3; tail -3 gpurun_out/r3c_bench.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r3c_bench.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["e2e"], d["host_enqueue_ms_per_step"], d["gpu_launches"], d["config"].get("cuda_graph"))
print({k: (v["ms_per_step"], v["launches_per_step"]) for k, v in d["kernel_breakdown"].items()})
print("roofline", {k: d["roofline"][k] for k in ("achieved", "frac", "frac_of_own_roofline", "tensor_bound_launches", "hbm_bound_launches", "gemm_ms_per_step")})
for k in ("retrieval", "sub_benches", "gpu_eager_reference", "cpu_baseline", "hbm_bound_kernels"):
    print(k, json.dumps(d.get(k))[:1500])
PY

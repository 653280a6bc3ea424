#!/bin/bash
# round 2, call P: full GPU suite, default bench (all legs), ncu launch list of one step + GEMM descriptor log
cd /root/repo
timeout 1800 python -m pytest tests -m gpu -q -x > gpurun_out/r2p_tests.log 2>&1; echo "gpu tests rc=$?"; tail -4 gpurun_out/r2p_tests.log
( time timeout 900 python bench.py > gpurun_out/r2p_bench.json 2> gpurun_out/r2p_bench.err ) 2>&1 | tail -3
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2p_bench.json").read().strip().splitlines()[-1])
print("train", d["value"], d["ms_per_step"], d["e2e"]["value"], d["host_enqueue_ms_per_step"], d["gpu_launches"])
print("gemm", d["kernel_breakdown"]["gemm"], d["roofline"]["frac"], d["roofline"]["frac_of_own_roofline"], d["roofline"]["tensor_bound_launches"])
r = d["retrieval"]; print("retr", r["value"], r.get("e2e"), r.get("cpu_baseline"))
print("sub", {k: (v.get("value"), v.get("e2e", {}).get("value")) for k, v in (d.get("sub_benches") or {}).items()})
print("eager", d.get("gpu_eager_reference")); print("cpu", d.get("cpu_baseline")); print("hbm", d.get("hbm_bound_kernels"))
PY
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r2p_launches_step.csv python tools/profile_step.py --dump-gemms > gpurun_out/r2p_profile_step.log 2>&1; echo "ncu list rc=$?"
cp gpurun_out/gemm_desc_log.json gpurun_out/r2p_gemm_desc_log.json 2>/dev/null

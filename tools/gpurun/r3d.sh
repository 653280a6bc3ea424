#!/bin/bash
# round 2 (session 2), call D: patchify_nchw fast path tests; ncu launch lists of ONE CUDA-graph replay of the training iteration and of
# one 808-pair retrieval forward; default bench
cd /root/repo
timeout 300 python -m pytest tests/test_kernels_gpu.py -q -x -k "patchify" > gpurun_out/r3d_tests.log 2>&1; echo "patchify tests rc=$?"; tail -3 gpurun_out/r3d_tests.log
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r3_launches_step.csv python tools/profile_step.py --graph > gpurun_out/r3d_profile_step.log 2>&1; echo "ncu list rc=$?"; tail -2 gpurun_out/r3d_profile_step.log
cp gpurun_out/gemm_desc_log.json gpurun_out/r3_gemm_desc_log.json
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r3_launches_retrieval.csv python tools/profile_step.py --retrieval --batch 808 > gpurun_out/r3d_profile_retr.log 2>&1; echo "ncu retrieval rc=$?"
( time timeout 900 python bench.py > gpurun_out/r3d_bench.json 2> gpurun_out/r3d_bench.err ) 2>&1 | tail -3; tail -3 gpurun_out/r3d_bench.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r3d_bench.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["e2e"], d["host_enqueue_ms_per_step"], d["gpu_launches"], d["config"].get("cuda_graph"))
print({k: (v["ms_per_step"], v["launches_per_step"]) for k, v in d["kernel_breakdown"].items() if "patch" in k})
r = d["retrieval"]; print("retr", r["value"], r["e2e"]["value"], {k: v["ms_per_step"] for k, v in r["kernel_breakdown"].items()})
print("sub", {k: (v.get("value"), v.get("e2e", {}).get("value")) for k, v in (d.get("sub_benches") or {}).items()})
PY

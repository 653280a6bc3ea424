#!/bin/bash
# round 2 (session 3), call A: FINAL validation of the tree -- full GPU suite (incl. the planted recognition protocol, the mlm_count
# claim test and the fused gradient clipping), smoke, default bench (all legs), ncu launch lists of ONE CUDA-graph replay of the
# training iteration in its final form (parallel branches, LayerNorm fusion, patch views) and of one 808-pair retrieval forward
cd /root/repo
S=$SECONDS
timeout 240 python -m pytest tests -m gpu -q > gpurun_out/r4a_tests.log 2>&1; echo "gpu tests rc=$? t=$((SECONDS-S))"; tail -4 gpurun_out/r4a_tests.log | cut -c1-300
grep -E "^(FAILED|ERROR)" gpurun_out/r4a_tests.log | cut -c1-250 | head -20
timeout 90 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1; echo "smoke t=$((SECONDS-S))"
timeout 200 python bench.py > gpurun_out/r4a_bench.json 2> gpurun_out/r4a_bench.err; echo "bench rc=$? t=$((SECONDS-S))"; tail -2 gpurun_out/r4a_bench.err | cut -c1-300
if [ $((SECONDS-S)) -lt 330 ]; then
  timeout 110 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r4_launches_step.csv python tools/profile_step.py --graph > gpurun_out/r4a_profile_step.log 2>&1; echo "ncu list rc=$? t=$((SECONDS-S))"; tail -1 gpurun_out/r4a_profile_step.log | cut -c1-200
  cp gpurun_out/gemm_desc_log.json gpurun_out/r4_gemm_desc_log.json 2>/dev/null
fi
if [ $((SECONDS-S)) -lt 400 ]; then
  timeout 70 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r4_launches_retrieval.csv python tools/profile_step.py --retrieval --batch 808 > gpurun_out/r4a_profile_retr.log 2>&1; echo "ncu retrieval rc=$? t=$((SECONDS-S))"
fi
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r4a_bench.json").read().strip().splitlines()[-1])
    print("train", d["value"], d["ms_per_step"], d["e2e"], d["host_enqueue_ms_per_step"], d["gpu_launches"], d["clocks"], d["config"].get("cuda_graph"))
    print("roofline", {k: d["roofline"].get(k) for k in ("achieved", "frac", "traffic", "frac_of_own_roofline", "gemm_ms_per_step")})
    r = d["retrieval"]; print("retr", r["value"], r.get("e2e", {}).get("value"))
    print("sub", {k: (v.get("value"), v.get("ms_per_step"), v.get("e2e", {}).get("value")) for k, v in (d.get("sub_benches") or {}).items()})
    print("eager", d.get("gpu_eager_reference", {}).get("value"), d.get("gpu_eager_reference", {}).get("ours_over_eager")); print("cpu", d.get("cpu_baseline", {}).get("value"))
    print("hbm", {k: v["frac_of_hbm_peak"] for k, v in d.get("hbm_bound_kernels", {}).items()})
except Exception as e:
    print("bench summary ERR", e)
for f in ("planted_recognition.json",):
    try:
        rows = json.load(open("gpurun_out/" + f))
        print(f, "n", len(rows), "min margin", min(r["margin"] for r in rows), "max err", max(r["max_err"] for r in rows),
              "mismatch", sum(r["pred_gpu"] != r["pred_oracle"] for r in rows))
    except Exception as e:
        print(f, "ERR", e)
PY
echo "total t=$((SECONDS-S))"

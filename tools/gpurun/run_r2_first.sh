#!/bin/bash
# PREPARED for the first GPU call of the next round (not run yet): validates the experimental fused attention backward
# (csrc/attn_bwd_tcgen05.cu) in isolation under a timeout, then through the model parity tests, then A/B in the bench.
#   /usr/local/graft/bin/gpurun --timeout 900 -- 'bash tools/gpurun/run_r2_first.sh'
cd /root/repo
python -m pytest tests -m gpu -q -x > gpurun_out/r2_tests_default.log 2>&1; echo "default tests rc=$?"; tail -3 gpurun_out/r2_tests_default.log
MVLT_FUSED_ATTN_BWD=1 timeout 180 python -m pytest tests/test_attention_gpu.py -q -x -k backward > gpurun_out/r2_attn_bwd.log 2>&1
rc=$?; echo "fused attention backward kernel tests rc=$rc"; tail -15 gpurun_out/r2_attn_bwd.log
if [ $rc -eq 0 ]; then
  MVLT_FUSED_ATTN_BWD=1 timeout 300 python -m pytest tests/test_model_gpu.py tests/test_engine_gpu.py -q -x > gpurun_out/r2_model_bwd.log 2>&1
  echo "model tests with the fused backward rc=$?"; tail -4 gpurun_out/r2_model_bwd.log
  MVLT_FUSED_ATTN_BWD=1 timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu --retrieval-queries 0 > gpurun_out/r2_bench_bwd1.json 2> gpurun_out/r2_bench_bwd1.err
fi
timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu --retrieval-queries 0 > gpurun_out/r2_bench_bwd0.json 2> gpurun_out/r2_bench_bwd0.err
python - <<'PY'
import json
for n in ("bwd0", "bwd1"):
    try:
        d = json.loads(open(f"gpurun_out/r2_bench_{n}.json").read().strip().splitlines()[-1])
        print(n, d["value"], d["ms_per_step"], d["e2e"]["value"], d.get("hbm_bound_kernels"))
    except Exception as e:
        print(n, "ERR", e)
PY
# retrieval step profile (never captured in round 1): per-kernel time of one ITM-only forward at the sweep's batch
timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r2_launches_retrieval.csv python tools/profile_step.py --retrieval --batch 808 > gpurun_out/r2_ncu_retr.log 2>&1; echo "ncu retrieval rc=$?"

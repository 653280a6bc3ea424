#!/bin/bash
# call B: attention v2 validation + A/B against v1 + profiles
cd /root/repo
timeout 240 python -m pytest tests/test_attention_gpu.py -x -q > gpurun_out/b_attn.log 2>&1
rc=$?
echo "attention v2 tests rc=$rc"; tail -12 gpurun_out/b_attn.log
if [ $rc -ne 0 ]; then
  export MVLT_ATTN_V1=1; echo "V2 FAILED -> V1 FOR THE REST"
  timeout 240 python -m pytest tests/test_attention_gpu.py -x -q 2>&1 | tail -3
fi
timeout 400 python -m pytest tests -m gpu -q --deselect tests/test_attention_gpu.py > gpurun_out/b_all.log 2>&1; echo "all tests rc=$?"; tail -4 gpurun_out/b_all.log
timeout 120 python tools/attn_sweep.py --fused 2>&1 | grep fused > gpurun_out/b_sweep_v2.log; echo "v2 (or v1 if failed):"; cat gpurun_out/b_sweep_v2.log
MVLT_ATTN_V1=1 timeout 120 python tools/attn_sweep.py --fused 2>&1 | grep fused > gpurun_out/b_sweep_v1.log; echo "v1:"; cat gpurun_out/b_sweep_v1.log
timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/b_bench.json 2>gpurun_out/b_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/b_bench.json").read().strip().splitlines()[-1])
    print("bench", d["value"], d["ms_per_step"], d["e2e"]["value"], d["retrieval"]["value"], d["roofline"]["frac_of_own_roofline"])
    print({k: v["ms_per_step"] for k, v in d["kernel_breakdown"].items()})
except Exception as e:
    print("ERR", e); print(open("gpurun_out/b_bench.err").read()[-1500:])
PY
timeout 400 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r1h_launches_step.csv python tools/profile_step.py > gpurun_out/b_ncu_list.log 2>&1; echo "ncu list rc=$?"
cp gpurun_out/gemm_desc_log.json gpurun_out/r1h_gemm_desc_log.json 2>/dev/null
timeout 300 ncu --set full --clock-control none --import-source on -k regex:sr_attention -s 2 -c 2 -o gpurun_out/r1h_attn python tools/attn_one.py > gpurun_out/b_ncu_attn.log 2>&1; echo "ncu attn rc=$?"
ls -la gpurun_out | tail -8

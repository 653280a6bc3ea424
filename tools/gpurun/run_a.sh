#!/bin/bash
# call A: fused attention validation + A/B + c4/c5 lines
cd /root/repo
timeout 240 python -m pytest tests/test_attention_gpu.py -x -q > gpurun_out/a_attn.log 2>&1
rc=$?
echo "attention tests rc=$rc"; tail -15 gpurun_out/a_attn.log
if [ $rc -ne 0 ]; then export MVLT_FUSED_ATTN=0; echo "FUSED ATTENTION DISABLED FOR THE REST"; fi
timeout 400 python -m pytest tests -m gpu -q --deselect tests/test_attention_gpu.py > gpurun_out/a_all.log 2>&1; echo "all tests rc=$?"; tail -8 gpurun_out/a_all.log
if [ $rc -eq 0 ]; then timeout 120 python tools/attn_sweep.py --fused > gpurun_out/a_sweep.log 2>&1; cat gpurun_out/a_sweep.log; fi
timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/a_bench_fused.json 2>gpurun_out/a_bench_fused.err; echo "bench rc=$?"
MVLT_FUSED_ATTN=0 timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu --retrieval-queries 200 > gpurun_out/a_bench_unfused.json 2>gpurun_out/a_bench_unfused.err
timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu --retrieval-queries 0 --model pvlt_small > gpurun_out/a_bench_small.json 2>gpurun_out/a_bench_small.err; echo "small rc=$?"
timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu --retrieval-queries 0 --workload recognition > gpurun_out/a_bench_recog.json 2>gpurun_out/a_bench_recog.err; echo "recog rc=$?"
python - <<'PY'
import json
for n in ("fused","unfused","small","recog"):
    try:
        d=json.loads(open(f"gpurun_out/a_bench_{n}.json").read().strip().splitlines()[-1])
        print(n, d["value"], d["ms_per_step"], d["e2e"]["value"], (d.get("retrieval") or {}).get("value"), d["roofline"]["frac_of_own_roofline"] if d.get("roofline") else None)
    except Exception as e:
        print(n, "ERR", e); print(open(f"gpurun_out/a_bench_{n}.err").read()[-1500:])
PY

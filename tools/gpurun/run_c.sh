#!/bin/bash
# call C: t2i loss rewrite + PDL on/off
cd /root/repo
MVLT_PDL=0 timeout 300 python -m pytest tests -m gpu -q -x > gpurun_out/c_tests_pdl0.log 2>&1; echo "tests PDL=0 rc=$?"; tail -4 gpurun_out/c_tests_pdl0.log
MVLT_PDL=1 timeout 300 python -m pytest tests -m gpu -q -x > gpurun_out/c_tests_pdl1.log 2>&1; echo "tests PDL=1 rc=$?"; tail -4 gpurun_out/c_tests_pdl1.log
MVLT_PDL=0 timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu --retrieval-queries 200 > gpurun_out/c_bench_pdl0.json 2>gpurun_out/c_bench_pdl0.err
MVLT_PDL=1 timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu --retrieval-queries 200 > gpurun_out/c_bench_pdl1.json 2>gpurun_out/c_bench_pdl1.err
MVLT_PDL=0 timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu --retrieval-queries 0 > gpurun_out/c_bench_pdl0b.json 2>gpurun_out/c_bench_pdl0b.err
python - <<'PY'
import json
for n in ("pdl0","pdl1","pdl0b"):
    try:
        d=json.loads(open(f"gpurun_out/c_bench_{n}.json").read().strip().splitlines()[-1])
        print(n, d["value"], d["ms_per_step"], d["e2e"]["value"], (d.get("retrieval") or {}).get("value"), d["host_enqueue_ms_per_step"])
        print({k: v["ms_per_step"] for k, v in d["kernel_breakdown"].items()})
    except Exception as e:
        print(n, "ERR", e); print(open(f"gpurun_out/c_bench_{n}.err").read()[-1500:])
PY

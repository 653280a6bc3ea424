#!/bin/bash
# round 2, call C: planted-positive retrieval protocol tuning (fit on GPU, rank vs fp32 oracle), weight-copy test, full default bench
cd /root/repo
timeout 600 python tools/planted_retrieval.py --steps 300 --queries 4 > gpurun_out/r2c_planted.log 2>&1; echo "planted rc=$?"; grep -v "^planted ITM fit step [0-9]*[1-9]:" gpurun_out/r2c_planted.log | tail -30
timeout 300 python -m pytest tests/test_engine_gpu.py tests/test_optim_gpu.py -q -x -k "adamw" > gpurun_out/r2c_tests.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/r2c_tests.log
( time timeout 900 python bench.py > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err ) 2>&1 | tail -3; tail -5 gpurun_out/r2c_bench.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2c_bench.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["host_enqueue_ms_per_step"], d["gpu_launches"])
print({k: (v["ms_per_step"], v["launches_per_step"]) for k, v in d["kernel_breakdown"].items()})
for k in ("retrieval", "sub_benches", "gpu_eager_reference", "cpu_baseline"):
    print(k, d.get(k))
print(d["roofline"])
PY
( time timeout 600 python bench.py --impl reference --steps 10 --warmup 3 ) 2>&1 | tail -5

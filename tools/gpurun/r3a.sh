#!/bin/bash
# round 2 (session 2), call A: CUDA-graph replay of the training iteration: unit / parity tests, then graph vs per-launch bench A/B
cd /root/repo
timeout 600 python -m pytest tests/test_graph_gpu.py -q -x > gpurun_out/r3a_tests.log 2>&1; echo "graph tests rc=$?"; tail -25 gpurun_out/r3a_tests.log
timeout 400 python bench.py --steps 20 --warmup 5 --no-cpu --no-eager --no-sub --retrieval-queries 0 > gpurun_out/r3a_bench_graph.json 2> gpurun_out/r3a_bench_graph.err; echo "bench graph rc=$?"; tail -5 gpurun_out/r3a_bench_graph.err
timeout 400 python bench.py --steps 20 --warmup 5 --no-cpu --no-eager --no-sub --retrieval-queries 0 --no-graph > gpurun_out/r3a_bench_nograph.json 2> gpurun_out/r3a_bench_nograph.err; echo "bench nograph rc=$?"; tail -5 gpurun_out/r3a_bench_nograph.err
python - <<'PY'
import json
for n in ("graph", "nograph"):
    try:
        d = json.loads(open(f"gpurun_out/r3a_bench_{n}.json").read().strip().splitlines()[-1])
        print(n, d["value"], d["ms_per_step"], d["e2e"], d["host_enqueue_ms_per_step"], d["gpu_launches"], d["config"].get("cuda_graph"))
    except Exception as e:
        print(n, "ERR", e)
PY

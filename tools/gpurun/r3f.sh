#!/bin/bash
# round 2 (session 2), call F: weight-gradient launches as a parallel branch of the captured graph: tests, then A/B in the bench
cd /root/repo
timeout 900 python -m pytest tests/test_graph_gpu.py tests/test_engine_gpu.py -q -x -k "graph or train_and_eval" > gpurun_out/r3f_tests.log 2>&1; echo "graph tests rc=$?"; tail -6 gpurun_out/r3f_tests.log
timeout 400 python bench.py --steps 20 --warmup 5 --no-cpu --no-eager --no-sub --retrieval-queries 0 > gpurun_out/r3f_bench_branch1.json 2> gpurun_out/r3f_bench_branch1.err; echo "bench branch rc=$?"; tail -3 gpurun_out/r3f_bench_branch1.err
MVLT_GRAPH_WGRAD_BRANCH=0 timeout 400 python bench.py --steps 20 --warmup 5 --no-cpu --no-eager --no-sub --retrieval-queries 0 > gpurun_out/r3f_bench_branch0.json 2> gpurun_out/r3f_bench_branch0.err; echo "bench no-branch rc=$?"
python - <<'PY'
import json
for n in ("branch1", "branch0"):
    try:
        d = json.loads(open(f"gpurun_out/r3f_bench_{n}.json").read().strip().splitlines()[-1])
        print(n, d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["host_enqueue_ms_per_step"], d["config"].get("cuda_graph"))
    except Exception as e:
        print(n, "ERR", e)
PY

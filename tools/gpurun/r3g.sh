#!/bin/bash
# round 2 (session 2), call G: more parallel branches in the captured graph (SR key/value chain, t2i head, pos-embed / decoder gradients on the
# weight-gradient stream): tests, then A/B of the number of branches
cd /root/repo
timeout 900 python -m pytest tests/test_graph_gpu.py tests/test_engine_gpu.py -q -x -k "graph or train_and_eval" > gpurun_out/r3g_tests.log 2>&1; echo "graph tests rc=$?"; tail -6 gpurun_out/r3g_tests.log
for nb in 2 1 0; do
MVLT_GRAPH_BRANCHES=$nb timeout 400 python bench.py --steps 20 --warmup 5 --no-cpu --no-eager --no-sub --retrieval-queries 0 > gpurun_out/r3g_bench_nb$nb.json 2> gpurun_out/r3g_bench_nb$nb.err; echo "bench nb=$nb rc=$?"; tail -2 gpurun_out/r3g_bench_nb$nb.err
done
python - <<'PY'
import json
for n in ("nb2", "nb1", "nb0"):
    try:
        d = json.loads(open(f"gpurun_out/r3g_bench_{n}.json").read().strip().splitlines()[-1])
        print(n, d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["host_enqueue_ms_per_step"])
    except Exception as e:
        print(n, "ERR", e)
PY

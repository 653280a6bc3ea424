#!/bin/bash
# round 2 (session 2), call I: backward spatial-reduction chain as a branch next to the query GEMM: tests + bench
cd /root/repo
timeout 900 python -m pytest tests/test_graph_gpu.py tests/test_kernels_gpu.py tests/test_engine_gpu.py -q -x -k "graph or train_and_eval or patchify or colsum" > gpurun_out/r3i_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/r3i_tests.log
timeout 400 python bench.py --steps 20 --warmup 5 --no-cpu --no-eager --no-sub --retrieval-queries 0 > gpurun_out/r3i_bench.json 2> gpurun_out/r3i_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/r3i_bench.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r3i_bench.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["host_enqueue_ms_per_step"])
PY

#!/bin/bash
# round 2 (session 2), call B (2 GPUs): graph tests, the NCCL gradient exchange captured inside the CUDA graph (dist_check), N=2 bench graph vs per-launch
cd /root/repo
timeout 600 python -m pytest tests/test_graph_gpu.py -q -x > gpurun_out/r3b_tests.log 2>&1; echo "graph tests rc=$?"; tail -5 gpurun_out/r3b_tests.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/dist_check.py > gpurun_out/r3b_dist_check.log 2>&1; echo "dist_check rc=$?"; tail -12 gpurun_out/r3b_dist_check.log
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 --no-sub --retrieval-queries 0 > gpurun_out/r3b_bench_n2_graph.json 2> gpurun_out/r3b_bench_n2_graph.err; echo "bench n2 graph rc=$?"; tail -3 gpurun_out/r3b_bench_n2_graph.err
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 20 --warmup 5 --no-sub --retrieval-queries 0 --no-graph > gpurun_out/r3b_bench_n2_nograph.json 2> gpurun_out/r3b_bench_n2_nograph.err; echo "bench n2 nograph rc=$?"; tail -3 gpurun_out/r3b_bench_n2_nograph.err
python - <<'PY'
import json
for n in ("graph", "nograph"):
    try:
        d = json.loads(open(f"gpurun_out/r3b_bench_n2_{n}.json").read().strip().splitlines()[-1])
        print(n, d["value"], d["ms_per_step"], d["e2e"], d["host_enqueue_ms_per_step"], d["gpu_launches"], d["config"].get("cuda_graph"))
    except Exception as e:
        print(n, "ERR", e)
PY

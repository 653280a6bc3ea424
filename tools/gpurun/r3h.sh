#!/bin/bash
# round 2 (session 2), call H (2 GPUs): multi-branch graph + NCCL exchange: dist_check, N=2 bench
cd /root/repo
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/dist_check.py > gpurun_out/r3h_dist_check.log 2>&1; echo "dist_check rc=$?"; grep -v "^\s*File\|^\s*\^\|OMP_NUM\|\*\*\*" gpurun_out/r3h_dist_check.log | tail -8
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 --no-sub --retrieval-queries 0 > gpurun_out/r3h_bench_n2.json 2> gpurun_out/r3h_bench_n2.err; echo "bench n2 rc=$?"; tail -3 gpurun_out/r3h_bench_n2.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r3h_bench_n2.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["e2e"], d["host_enqueue_ms_per_step"], d["config"].get("cuda_graph"))
PY

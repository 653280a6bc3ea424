#!/bin/bash
# round 2, call Q (2 GPUs): overlapped gradient exchange: consistency check, N=2 bench; N=1 bench on the same box for the ratio
cd /root/repo
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/dist_check.py 2>&1 | grep -v "^W\|OMP_NUM" | tail -8
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 --no-sub --retrieval-queries 200 > gpurun_out/r2q_bench_n2.json 2> gpurun_out/r2q_bench_n2.err; tail -3 gpurun_out/r2q_bench_n2.err
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu --no-eager --no-sub --retrieval-queries 200 > gpurun_out/r2q_bench_n1.json 2> gpurun_out/r2q_bench_n1.err
python - <<'PY'
import json
for n in ("n1", "n2"):
    try:
        d = json.loads(open(f"gpurun_out/r2q_bench_{n}.json").read().strip().splitlines()[-1])
        print(n, d["value"], d["ms_per_step"], d["e2e"]["value"], d["retrieval"]["value"])
    except Exception as e:
        print(n, "ERR", e)
PY

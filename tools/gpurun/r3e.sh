#!/bin/bash
# round 2 (session 2), call E (8 GPUs): the driver's SCALE command at N = 8 with CUDA-graph replay (default) and the per-launch A/B
cd /root/repo
( time timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r3e_bench_n8.json 2> gpurun_out/r3e_bench_n8.err ) 2>&1 | tail -3; echo "ours rc=$?"; tail -5 gpurun_out/r3e_bench_n8.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 10 --warmup 3 --no-graph --no-sub --retrieval-queries 0 > gpurun_out/r3e_bench_n8_nograph.json 2> gpurun_out/r3e_bench_n8_nograph.err; echo "nograph rc=$?"
python - <<'PY'
import json
for n in ("r3e_bench_n8", "r3e_bench_n8_nograph"):
    try:
        d = json.loads(open(f"gpurun_out/{n}.json").read().strip().splitlines()[-1])
        print(n, "train", d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], "host", d.get("host_enqueue_ms_per_step"), d["config"].get("cuda_graph"))
        r = d.get("retrieval") or {}; print(" retr", r.get("value"), r.get("e2e"), r.get("error"))
        print(" sub", {k: (v.get("value"), v.get("ms_per_step"), v.get("e2e", {}).get("value"), v.get("error")) for k, v in (d.get("sub_benches") or {}).items()})
        print(" clocks", d.get("clocks"))
    except Exception as e:
        print(n, "ERR", e)
PY

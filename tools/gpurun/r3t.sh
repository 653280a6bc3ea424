#!/bin/bash
# round 2 (session 2), call T (8 GPUs, FINAL): the driver's SCALE command at N = 8, final state (graph replay with parallel branches)
cd /root/repo
( time timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r3t_bench_n8.json 2> gpurun_out/r3t_bench_n8.err ) 2>&1 | tail -3; echo "ours rc=$?"; tail -3 gpurun_out/r3t_bench_n8.err
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-eager --no-sub --retrieval-queries 0 > gpurun_out/r3t_bench_n1_same_box.json 2> gpurun_out/r3t_bench_n1.err; echo "n1 rc=$?"
python - <<'PY'
import json
for n in ("r3t_bench_n8", "r3t_bench_n1_same_box"):
    try:
        d = json.loads(open(f"gpurun_out/{n}.json").read().strip().splitlines()[-1])
        print(n, "train", d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], "host", d.get("host_enqueue_ms_per_step"))
        r = d.get("retrieval") or {}; print(" retr", r.get("value"), (r.get("e2e") or {}).get("value"), r.get("error"))
        print(" sub", {k: (v.get("value"), v.get("ms_per_step"), v.get("e2e", {}).get("value"), v.get("error")) for k, v in (d.get("sub_benches") or {}).items()})
        print(" clocks", d.get("clocks"))
    except Exception as e:
        print(n, "ERR", e)
PY

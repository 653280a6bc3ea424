#!/bin/bash
# round 2, call X (8 GPUs): the driver's SCALE run at N = 8, both arms, as it will be launched at round end
cd /root/repo
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --impl reference --gpus 8 --steps 10 --warmup 3 > gpurun_out/r2x_ref_n8.json 2> gpurun_out/r2x_ref_n8.err; echo "reference arm rc=$?"; tail -c 400 gpurun_out/r2x_ref_n8.json
( time timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r2x_bench_n8.json 2> gpurun_out/r2x_bench_n8.err ) 2>&1 | tail -3; echo "ours rc=$?"; tail -5 gpurun_out/r2x_bench_n8.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2x_bench_n8.json").read().strip().splitlines()[-1])
print("N=8 train", d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"])
r = d["retrieval"]; print("retr", r.get("value"), r.get("e2e"), r.get("error"))
print("sub", {k: (v.get("value"), v.get("ms_per_step"), v.get("error")) for k, v in (d.get("sub_benches") or {}).items()})
print("clocks", d.get("clocks"))
PY

#!/bin/bash
# round 2, call Y (8 GPUs): A/B of the gradient exchange at N = 8: overlapped segments vs one exchange after the backward, NCCL CTA cap
cd /root/repo
run() {
  tag=$1; shift
  env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 20 --warmup 5 --no-sub --retrieval-queries 0 > gpurun_out/r2y_$tag.json 2> gpurun_out/r2y_$tag.err
  python - "$tag" <<'PY'
import json, sys
try:
    d = json.loads(open(f"gpurun_out/r2y_{sys.argv[1]}.json").read().strip().splitlines()[-1])
    print(sys.argv[1], d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"])
except Exception as e:
    print(sys.argv[1], "ERR", e)
PY
}
run overlap1 MVLT_GRAD_OVERLAP=1
run overlap0 MVLT_GRAD_OVERLAP=0
run overlap1_cta8 MVLT_GRAD_OVERLAP=1 NCCL_MAX_CTAS=8
run overlap0_cta16 MVLT_GRAD_OVERLAP=0 NCCL_MAX_CTAS=16
run overlap1_b MVLT_GRAD_OVERLAP=1
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu --no-eager --no-sub --retrieval-queries 0 > gpurun_out/r2y_n1.json 2>/dev/null
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2y_n1.json").read().strip().splitlines()[-1])
print("n1", d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"])
PY

#!/bin/bash
# round 2 (session 2), call J: retrieval sweep with / without the key/value branch stream
cd /root/repo
for nb in 0 1; do
MVLT_RETR_BRANCHES=$nb timeout 300 python - <<'PY'
import os, torch
from mvlt_b200 import retrieval
dev = torch.device("cuda", 0)
r = retrieval.bench_sweep(dev, 0, 1, n_query=1000, n_cand=101, warmup=1)
print("branches", os.environ["MVLT_RETR_BRANCHES"], r["value"], r["e2e"]["value"])
PY
done

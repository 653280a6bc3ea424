import os, sys, torch
sys.path.insert(0, os.getcwd())
from mvlt_b200 import retrieval
dev = torch.device("cuda", 0)
for q in (2, 4, 8, 16):
    r = retrieval.bench_sweep(dev, 0, 1, n_query=192, n_cand=101, queries_per_step=q, warmup=1)
    print(q, r["value"], r["ms_total"], flush=True)

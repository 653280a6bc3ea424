#!/bin/bash
# round 2, call K: where does the fused MLP forward spend its time? (debug switches: residual read / GELU math / output path)
cd /root/repo
for d in 0 1 2 4 6 7; do echo "MVLT_MLP_DBG=$d"; MVLT_MLP_DBG=$d timeout 120 python tools/mlp_bench.py --iters 5 2>&1 | grep fused; done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:mlp_fwd -s 2 -c 1 -f -o gpurun_out/r2k_mlp_fwd python tools/mlp_bench.py --iters 1 > gpurun_out/r2k_ncu.log 2>&1; echo "ncu rc=$?"

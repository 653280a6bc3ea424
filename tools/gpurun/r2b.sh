#!/bin/bash
# round 2, call B: full GPU suite incl. the new parity tests (B=128, drop-path/dropout on, weight-copy freshness), bench
cd /root/repo
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r2b_tests.log 2>&1; echo "gpu tests rc=$?"; tail -25 gpurun_out/r2b_tests.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu --retrieval-queries 0 > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err; tail -c 1500 gpurun_out/r2b_bench.json
free -g | head -2; nproc

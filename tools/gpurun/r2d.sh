#!/bin/bash
# round 2, call D: which planted task does an ITM-only PVLT-tiny fit quickly? (classes x lr grid, no oracle)
cd /root/repo
for cfg in "2 5e-4" "2 2e-3" "8 5e-4" "8 2e-3" "64 1e-3"; do
  set -- $cfg
  echo "=== classes $1 lr $2"
  timeout 300 python tools/planted_retrieval.py --steps 1200 --batch 128 --classes $1 --lr $2 --queries 2 --oracle 0 2>&1 | grep -E "step (0|100|200|300|400|500|600|700|800|900|1000|1100|1199):|rank_gpu|fit time|Error|error" 
done

#!/bin/bash
# round 2, call R: ncu launch list of one step (+ GEMM descriptor log), planted retrieval test, retrieval-forward launch list
cd /root/repo
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r2r_launches_step.csv python tools/profile_step.py > gpurun_out/r2r_profile_step.log 2>&1; echo "ncu list rc=$?"
cp gpurun_out/gemm_desc_log.json gpurun_out/r2r_gemm_desc_log.json
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r2r_launches_retrieval.csv python tools/profile_step.py --retrieval --batch 808 > gpurun_out/r2r_profile_retr.log 2>&1; echo "ncu retrieval rc=$?"
timeout 900 python -m pytest tests/test_engine_gpu.py -q -x -k planted 2>&1 | tail -3

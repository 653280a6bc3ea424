#!/bin/bash
# round 2 (session 3), call B (the round's last GPU seconds): ncu --set full of the fused attention backward (stage-1 launches: the
# kernel DESIGN calls latency-bound) and compute-sanitizer memcheck / racecheck over the kernels added in this session (gradient-norm
# partial sums + clip coefficient, fixed-capacity compaction behind a caller-supplied mlm_count)
cd /root/repo
S=$SECONDS
timeout 90 ncu --set full --clock-control none --import-source on -k regex:sr_attention_bwd -s 6 -c 2 -f -o gpurun_out/r4b_attn_bwd python tools/profile_step.py > gpurun_out/r4b_ncu.log 2>&1; echo "ncu attn_bwd rc=$? t=$((SECONDS-S))"
timeout 80 compute-sanitizer --tool memcheck python -m pytest tests/test_optim_gpu.py "tests/test_model_gpu.py::test_supplied_mlm_count_is_a_claim_not_an_index_bound" -q > gpurun_out/r4b_sanitizer_memcheck.log 2>&1; echo "memcheck rc=$? t=$((SECONDS-S))"; grep -E "passed|failed|ERROR SUMMARY" gpurun_out/r4b_sanitizer_memcheck.log | tail -3
if [ $((SECONDS-S)) -lt 150 ]; then
  timeout 60 compute-sanitizer --tool racecheck python -m pytest tests/test_optim_gpu.py -q -k "clip" > gpurun_out/r4b_sanitizer_racecheck.log 2>&1; echo "racecheck rc=$? t=$((SECONDS-S))"; grep -E "passed|failed|RACECHECK SUMMARY" gpurun_out/r4b_sanitizer_racecheck.log | tail -3
fi
ls -la gpurun_out/r4b_attn_bwd.ncu-rep 2>/dev/null
echo "total t=$((SECONDS-S))"

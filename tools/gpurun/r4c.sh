#!/bin/bash
# round 2 (session 3), call C (the round's last GPU seconds): default bench of the final tree once more (AdamW byte accounting now counts
# the bf16 copies), the end-to-end example (main_vl.py flow on synthetic data: CUDA-graph replay + fused clipping; recognition task
# on the per-launch path + clipping; checkpoint round trip), then the optimizer / graph / loop tests while time remains
cd /root/repo
S=$SECONDS
timeout 100 python bench.py > gpurun_out/r4c_bench.json 2> gpurun_out/r4c_bench.err; echo "bench rc=$? t=$((SECONDS-S))"; tail -2 gpurun_out/r4c_bench.err | cut -c1-300
timeout 60 python examples/train_synthetic.py --epochs 2 --steps 4 --batch-size 16 --cuda-graph --clip-grad 1.0 --retrieval-queries 2 --checkpoint gpurun_out/r4c_ckpt.pth > gpurun_out/r4c_example_pretrain.log 2>&1; echo "example pretrain rc=$? t=$((SECONDS-S))"; grep -E "^epoch|^evaluate|^checkpoint|^done|Error|error" gpurun_out/r4c_example_pretrain.log | cut -c1-300 | tail -8
rm -f gpurun_out/r4c_ckpt.pth
if [ $((SECONDS-S)) -lt 125 ]; then
  timeout 45 python examples/train_synthetic.py --task recognition --epochs 2 --steps 4 --batch-size 16 --clip-grad 0.5 > gpurun_out/r4c_example_recog.log 2>&1; echo "example recognition rc=$? t=$((SECONDS-S))"; grep -E "^epoch|^evaluate|^done|Error|error" gpurun_out/r4c_example_recog.log | cut -c1-300 | tail -6
fi
if [ $((SECONDS-S)) -lt 130 ]; then
  timeout 40 python -m pytest tests/test_optim_gpu.py tests/test_graph_gpu.py -q -k "clip or adamw" > gpurun_out/r4c_tests.log 2>&1; echo "tests rc=$? t=$((SECONDS-S))"; tail -1 gpurun_out/r4c_tests.log
fi
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r4c_bench.json").read().strip().splitlines()[-1])
    print("train", d["value"], d["ms_per_step"], d["e2e"]["value"], d["gpu_launches"], d["clocks"])
    print("hbm", {k: v["frac_of_hbm_peak"] for k, v in d.get("hbm_bound_kernels", {}).items()})
    print("retr", d["retrieval"]["value"], "sub", {k: v.get("value") for k, v in (d.get("sub_benches") or {}).items()})
except Exception as e:
    print("bench summary ERR", e)
PY
echo "total t=$((SECONDS-S))"

"""Join the ncu launch list of one step with the GEMM descriptor log of the same step (same launch order):
python tools/join_gemm_log.py launches.csv gemm_desc_log.json  -> per-shape table with each shape's own roofline."""
import json
import os
import sys
from collections import defaultdict

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from launch_table import load  # noqa: E402

HBM, TF = 6456.2e9, 1419.1e12
L = [d for d in load(sys.argv[1]) if d["name"].startswith("gemm_tcgen05_kernel")]
G = json.load(open(sys.argv[2]))
assert len(L) == len(G), (len(L), len(G))
agg = defaultdict(lambda: [0, 0.0, 0.0, 0.0, 0.0, ""])
for d, (fl, by, what) in zip(L, G):
    a = agg[what + " | " + d["name"].replace("gemm_tcgen05_kernel", "k")]
    us = d["gpu__time_duration.sum"]
    a[0] += 1; a[1] += us; a[2] += max(by / HBM, fl / TF) * 1e6; a[3] += by; a[4] += fl
tot = sum(a[1] for a in agg.values()); ideal = sum(a[2] for a in agg.values())
print(f"{len(L)} GEMM launches: {tot / 1e3:.3f} ms measured (ncu, cold), {ideal / 1e3:.3f} ms at each launch's own roofline")
print(f"{'n':>3s} {'us':>8s} {'ideal':>7s} {'lost':>7s} {'GB/s':>6s} {'TF/s':>6s}  shape")
for k, a in sorted(agg.items(), key=lambda kv: -(kv[1][1] - kv[1][2])):
    print(f"{a[0]:3d} {a[1]:8.1f} {a[2]:7.1f} {a[1] - a[2]:7.1f} {a[3] / a[1] / 1e3:6.0f} {a[4] / a[1] / 1e6:6.0f}  {k}")

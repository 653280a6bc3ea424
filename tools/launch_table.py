"""Aggregate an ncu launch list (csv with gpu__time_duration.sum [+ dram bytes]) per kernel: python tools/launch_table.py csv [--md]"""
import csv
import re
import sys
from collections import defaultdict


def load(path):
    rows = list(csv.reader(open(path)))
    hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    h = rows[hi]
    iid, ik, im, iv, iu = h.index("ID"), h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value"), h.index("Metric Unit")
    launches = {}
    for r in rows[hi + 1:]:
        if len(r) < len(h):
            continue
        v = float(r[iv].replace(",", ""))
        u = r[iu]
        scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1.0)
        name = re.sub(r"<unnamed>::", "", r[ik])
        name = re.sub(r"\((?:[^()]|\([^()]*\))*\)$", "", name).replace("void ", "")
        d = launches.setdefault(int(r[iid]), {"name": name})
        d[r[im]] = v * scale
    return list(launches.values())


def main():
    L = load(sys.argv[1])
    if "--gemm-traffic" in sys.argv:      # dram bytes per launch of the dominant kernel, for bench.py's roofline.traffic
        import json
        G = [d for d in L if d["name"].startswith("gemm_tcgen05_kernel")]
        tb = sum(d.get("dram__bytes_read.sum", 0.0) + d.get("dram__bytes_write.sum", 0.0) for d in G)
        tt = sum(d.get("gpu__time_duration.sum", 0.0) for d in G)
        out = {"kernel": "gemm_tcgen05_kernel", "launches": len(G), "dram_bytes_per_launch": round(tb / max(len(G), 1)),
               "dram_bytes_total": round(tb), "gpu_time_us_total": round(tt, 1), "source": sys.argv[1],
               "note": "dram__bytes_read.sum + dram__bytes_write.sum per launch, averaged over all GEMM launches of one training step"}
        json.dump(out, open(sys.argv[sys.argv.index("--gemm-traffic") + 1], "w"), indent=1)
    agg = defaultdict(lambda: [0, 0.0, 0.0])
    for d in L:
        a = agg[d["name"]]
        a[0] += 1
        a[1] += d.get("gpu__time_duration.sum", 0.0)
        a[2] += d.get("dram__bytes_read.sum", 0.0) + d.get("dram__bytes_write.sum", 0.0)
    tot = sum(a[1] for a in agg.values())
    totb = sum(a[2] for a in agg.values())
    print(f"{len(L)} launches, {tot / 1e3:.3f} ms of kernel time, {totb / 1e9:.2f} GB of DRAM traffic "
          f"({totb / tot / 1e3 if tot else 0:.0f} GB/s average)")
    print(f"{'kernel':70s} {'n':>4s} {'us':>9s} {'share':>6s} {'GB':>7s} {'GB/s':>6s}")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k[:70]:70s} {a[0]:4d} {a[1]:9.1f} {100 * a[1] / tot:5.1f}% {a[2] / 1e9:7.3f} {a[2] / a[1] / 1e3 if a[1] else 0:6.0f}")


if __name__ == "__main__":
    main()

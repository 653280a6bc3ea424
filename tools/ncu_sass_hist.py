"""Opcode histogram + hottest SASS instructions of an .ncu-rep captured with --import-source on."""
import csv
import io
import subprocess
import sys
from collections import Counter

rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hi = next(i for i, r in enumerate(rows) if "Source" in r)
h = rows[hi]
isrc, iex = h.index("Source"), h.index("Instructions Executed")
ist = h.index("Warp Stall Sampling (All Samples)")
body = []
for r in rows[hi + 1:]:      # first kernel of the report only (a second launch repeats the header rows)
    if "Source" in r or (r and r[0] == "Kernel Name"):
        break
    if len(r) == len(h):
        body.append(r)
tot = sum(int(r[iex]) for r in body)
print("total warp instructions", tot)
c, s = Counter(), Counter()
for r in body:
    t = r[isrc].split()
    op = t[1] if t[0].startswith("@") else t[0]
    c[op.split(".")[0]] += int(r[iex])
    s[op.split(".")[0]] += int(r[ist])
stot = sum(s.values())
for k, v in c.most_common(28):
    print(f"{k:12s} {v:12d} {100 * v / tot:5.1f}%   stall samples {100 * s[k] / max(stot, 1):5.1f}%")
print("-- hottest by stall samples")
for r in sorted(body, key=lambda r: -int(r[ist]))[:14]:
    print(r[ist], r[iex], r[isrc].strip()[:110])

"""Summarise `ncu -i X.ncu-rep --page source --csv --print-source sass` (stdin or file): stall samples by reason, by SASS
mnemonic, and the top lines.  The report may hold several kernels (one CSV block each); the LAST block is summarised."""
import csv
import re
import sys
from collections import Counter

rows = list(csv.reader(open(sys.argv[1]) if len(sys.argv) > 1 else sys.stdin))
starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
block = rows[starts[-1]:]
print("kernel:", block[0][1][:150])
hdr = block[1]
body = [r for r in block[2:] if len(r) == len(hdr)]
col = {n: i for i, n in enumerate(hdr)}
reasons = [n for n in hdr if n.startswith("stall_") and "Not Issued" not in n]
num = lambda v: int(float(v)) if v not in ("", "-") else 0
tot = Counter()
by_op = Counter()
ex_op = Counter()
lines = []
for k, r in enumerate(body):
    s = num(r[col["# Samples"]])
    ex = num(r[col["Instructions Executed"]])
    src = r[col["Source"]]
    op = re.sub(r"^@!?U?P\d+\s+", "", src.strip()).split(" ")[0]
    rs = {n[6:]: num(r[col[n]]) for n in reasons}
    for n, v in rs.items():
        tot[n] += v
    by_op[op] += s
    ex_op[op] += ex
    lines.append((s, k, ex, src, rs))
total = sum(l[0] for l in lines)
print("rows", len(body), "samples", total, "warp instructions", sum(l[2] for l in lines))
print("by reason:", [(n, v, f"{100 * v / max(total, 1):.1f}%") for n, v in tot.most_common(12)])
print("by mnemonic (samples, share, executed):")
for op, s in by_op.most_common(25):
    print(f"  {op:28s} {s:7d} {100 * s / max(total, 1):5.1f}%  exec {ex_op[op]}")
print("top lines:")
for s, k, ex, src, rs in sorted(lines, reverse=True)[:int(sys.argv[2]) if len(sys.argv) > 2 else 30]:
    top = sorted(rs.items(), key=lambda kv: -kv[1])[:3]
    print(f"  {k:5d} {s:6d} {100 * s / max(total, 1):5.1f}% exec={ex:9d} {src[:90]:90s} {top}")

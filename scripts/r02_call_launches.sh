#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 420 --csv --log-file gpurun_out/r02_zd_launches_bf16.csv \
   python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-rollout --precision bf16 > gpurun_out/ncu_bench.log 2>&1
echo "launch list exit $?"
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open("gpurun_out/r02_zd_launches_bf16.csv")) if len(r) > 10]
hdr = rows[0]; ik, iv = hdr.index("Kernel Name"), hdr.index("Metric Value")
acc = collections.Counter(); n = collections.Counter()
for r in rows[1:]:
    name = r[ik]
    for key in ("OpConv", "OpIdft", "OpDft", "OpDhconv", "OpIleg", "OpLeg"):
        if key in name:
            name = key; break
    else:
        name = name.split("(")[0][:40]
    acc[name] += float(r[iv].replace(",", "")); n[name] += 1
tot = sum(acc.values())
print("total ns", tot, "launches", sum(n.values()))
for k, v in acc.most_common(10): print(f"{k:40s} {n[k]:4d} launches {100*v/tot:5.1f} %")
PY

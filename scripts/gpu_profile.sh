#!/bin/bash
# ncu evidence for profiles/: launch list of one bench step + full captures of the hot kernels (selftest shapes = ACE, batch 8).
# The .ncu-rep files are summarised ON THE BOX (gpurun brings back at most 64 MiB): per-kernel table, DRAM traffic per launch.
set -u
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 420 --csv --log-file gpurun_out/launches.csv \
   python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-rollout --precision bf16 > gpurun_out/ncu_bench.log 2>&1
echo "launch list exit $?"
CASES="fc1 6 8 256 512 64800 1 3
fc2 6 8 512 256 64800 0 5
skip 6 8 256 256 64800 1 1
idft 4 8 256 180 360 181 7
dft 0 8 256 180 360 181 0
leg 1 8 256 180 180 181 1
ileg 3 8 256 180 180 181 2
dhconv 2 8 256 180 181 1 0" bash scripts/ncu_selftest.sh > /dev/null
python scripts/summarize_ncu.py gpurun_out/ncu_kernels.md
python - <<'PY'
import csv, glob, io, json, os, subprocess
names = {"fc1": "mlp_fc1", "fc2": "mlp_fc2", "skip": "inner_skip", "idft": "dft_inv", "dft": "dft_fwd", "leg": "legendre_fwd",
         "ileg": "legendre_inv", "dhconv": "dhconv"}
out = {"source": "ncu --set full --clock-control none, one launch of each tensor-core kernel at ACE size (B = 8, bf16; triangular "
                 "Legendre / dhconv ranges as in the net) through tests/tc_selftest_cli.py (scripts/gpu_profile.sh); "
                 "dram__bytes_read.sum + dram__bytes_write.sum per launch", "kernels": {}}
for rep in sorted(glob.glob("gpurun_out/prof_*.ncu-rep")):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    if len(rows) < 3:
        continue
    hdr, units, r = rows[0], rows[1], rows[-1]
    def val(k):
        v, u = float(r[hdr.index(k)]), units[hdr.index(k)].lower()
        return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)
    rd, wr = val("dram__bytes_read.sum"), val("dram__bytes_write.sum")
    out["kernels"][names[os.path.basename(rep)[5:-8]]] = {"dram_read_bytes": rd, "dram_write_bytes": wr, "traffic_bytes": rd + wr}
json.dump(out, open("gpurun_out/ncu_traffic.json", "w"), indent=1)
print(json.dumps({k: round(v["traffic_bytes"] / 1e6, 1) for k, v in out["kernels"].items()}))
PY
rm -f gpurun_out/prof_*.ncu-rep
ls -la gpurun_out

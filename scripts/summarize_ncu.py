#!/usr/bin/env python
"""Summarise gpurun_out/prof_*.ncu-rep (ncu --set full captures) into a small markdown table for profiles/."""
import csv
import glob
import io
import os
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "time"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram %"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 %"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps %"),
    ("lts__t_sector_hit_rate.pct", "L2 hit %"),
    ("launch__registers_per_thread", "regs"),
    ("smsp__inst_executed.sum", "warp instr"),
]


def main(out_path):
    lines = ["| capture | kernel | " + " | ".join(k[1] for k in KEYS) + " |", "|---|---|" + "---|" * len(KEYS)]
    for rep in sorted(glob.glob("gpurun_out/prof_*.ncu-rep")):
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(raw)))
        if len(rows) < 3:
            continue
        hdr, units, r = rows[0], rows[1], rows[-1]
        name = r[hdr.index("Kernel Name")]
        name = name.replace("void sfno::gemm_tc_kernel<sfno::", "").split(">(")[0][:60]
        vals = []
        for k, _ in KEYS:
            if k in hdr:
                v, u = r[hdr.index(k)], units[hdr.index(k)]
                try:
                    v = f"{float(v):.4g}"
                except ValueError:
                    pass
                vals.append(f"{v} {u}".strip())
            else:
                vals.append("-")
        lines.append(f"| {os.path.basename(rep)[5:-8]} | `{name}` | " + " | ".join(vals) + " |")
    open(out_path, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "profiles/ncu_kernels.md")

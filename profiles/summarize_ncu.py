#!/usr/bin/env python
"""Turn the ncu artefacts a gpurun call left in gpurun_out/ into the tracked summaries under profiles/.

    python profiles/summarize_ncu.py <tag>      e.g. r01

Inputs  gpurun_out/launches_<tag>.csv            (ncu --metrics gpu__time_duration.sum launch list)
        gpurun_out/prof_<kernel>_<tag>.ncu-rep   (ncu --set full captures)
Outputs profiles/<tag>_launches_summary.md, profiles/<tag>_launches.csv.gz, profiles/<tag>_<kernel>_raw.csv
"""
import collections
import csv
import glob
import gzip
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
out_md = [f"# ncu summary {tag}\n"]

lpath = os.path.join(ROOT, "gpurun_out", f"launches_{tag}.csv")
if os.path.exists(lpath):
    rows = [r for r in csv.reader(open(lpath)) if len(r) > 10]
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    tot = collections.defaultdict(float); cnt = collections.Counter()
    for r in rows[1:]:
        name = r[ki].split("(")[0].replace("void ", "")
        v = float(r[vi].replace(",", ""))
        v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[ui], 1.0)
        tot[name] += v; cnt[name] += 1
    T = sum(tot.values())
    out_md.append(f"## launch list ({sum(cnt.values())} launches, `ncu --metrics gpu__time_duration.sum --clock-control none`)\n")
    out_md.append("Per-launch times are cold-cache and serialised: compare SHARES, not absolutes.\n")
    out_md.append("| kernel | launches | total ms | share |\n|---|---:|---:|---:|")
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
        out_md.append(f"| `{k}` | {cnt[k]} | {v:.3f} | {v / T:.4f} |")
    out_md.append(f"| total | {sum(cnt.values())} | {T:.3f} | 1 |\n")
    with open(lpath, "rb") as f, gzip.open(os.path.join(ROOT, "profiles", f"{tag}_launches.csv.gz"), "wb") as g:
        shutil.copyfileobj(f, g)

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__occupancy_limit_registers", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "l1tex__t_bytes.sum", "lts__t_bytes.sum"]
for rep in sorted(glob.glob(os.path.join(ROOT, "gpurun_out", f"prof_*_{tag}.ncu-rep"))):
    kname = os.path.basename(rep)[len("prof_"):-len(f"_{tag}.ncu-rep")]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    open(os.path.join(ROOT, "profiles", f"{tag}_{kname}_raw.csv"), "w").write(raw)
    rows = list(csv.reader(raw.splitlines()))
    if len(rows) < 3:
        continue
    hdr, units = rows[0], rows[1]
    out_md.append(f"## `{kname}` (`ncu --set full --clock-control none --import-source on`, {len(rows) - 2} launches)\n")
    out_md.append("| metric | unit | " + " | ".join(f"launch {i}" for i in range(len(rows) - 2)) + " |")
    out_md.append("|---|---|" + "---:|" * (len(rows) - 2))
    for m in WANT:
        if m in hdr:
            i = hdr.index(m)
            out_md.append(f"| {m} | {units[i]} | " + " | ".join(r[i] for r in rows[2:]) + " |")
    if "Kernel Name" in hdr:
        out_md.append(f"\nkernel: `{rows[2][hdr.index('Kernel Name')]}`\n")
open(os.path.join(ROOT, "profiles", f"{tag}_ncu_summary.md"), "w").write("\n".join(out_md) + "\n")
# machine-readable DRAM traffic per captured launch (bench.py copies it into roofline.traffic_ncu)
import json
traffic = {}
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0}
for rep in sorted(glob.glob(os.path.join(ROOT, "profiles", f"{tag}_*_raw.csv"))):
    kname = os.path.basename(rep)[len(tag) + 1:-len("_raw.csv")]
    rows = list(csv.reader(open(rep)))
    if len(rows) < 3:
        continue
    hdr, units, r = rows[0], rows[1], rows[2]
    def val(m):
        i = hdr.index(m); return float(r[i].replace(",", "")) * UNIT.get(units[i], 1.0)
    traffic[kname] = {"dram_bytes_read": val("dram__bytes_read.sum"), "dram_bytes_write": val("dram__bytes_write.sum"),
                      "duration_s": val("gpu__time_duration.sum"), "grid": r[hdr.index("launch__grid_size")],
                      "capture": ("ncu --set full --clock-control none, launch #101 of the class in bench.py (j = 101)"
                                  if kname in ("multidot", "axpy_dot", "multiaxpy", "multiaxpy_fin", "stencil", "stencil_smem")
                                  else "ncu --set full --clock-control none, profiles/ncu_targets.py " + kname)}
json.dump(traffic, open(os.path.join(ROOT, "profiles", f"{tag}_traffic.json"), "w"), indent=1)
print("\n".join(out_md))

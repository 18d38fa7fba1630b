#!/usr/bin/env python
"""Turns the raw ncu outputs a GPU run leaves in gpurun_out/ into the tracked summaries under profiles/.

  launch list   gpurun_out/launches_<tag>.csv  (ncu --metrics gpu__time_duration.sum --csv --log-file ...)
                -> profiles/<tag>_launches.csv (copy) + profiles/<tag>_launch_summary.txt (per-kernel totals, shares)
  full capture  gpurun_out/prof_<tag>.ncu-rep  (ncu --set full ...), read with `ncu -i ... --page raw --csv`
                -> profiles/<tag>_ncu_full_summary.{txt,json} (the metrics DESIGN.md / bench.py quote)
usage: python scripts/summarize_profiles.py r1_tensor ["header line for the summaries"]
"""
import csv
import json
import os
import re
import shutil
import subprocess
import sys
from collections import OrderedDict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEEP = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_subpipe_imma_cycles_active.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum", "lts__t_sector_hit_rate.pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__cluster_size",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__cycles_elapsed.avg.per_second",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__sass_inst_executed_op_utcmma.sum", "smsp__sass_inst_executed_op_tma_ld.sum",
    "smsp__sass_inst_executed_op_tmem_ldt.sum",
]


def short(name):
    name = re.sub(r"\(.*$", "", name)
    name = name.replace("<unnamed>::", "").replace("void ", "")
    return name.strip()


def launches(tag, header):
    src = os.path.join(ROOT, "gpurun_out", f"launches_{tag}.csv")
    if not os.path.exists(src):
        print("no", src)
        return
    shutil.copy(src, os.path.join(ROOT, "profiles", f"{tag}_launches.csv"))
    rows = [r for r in csv.reader(l for l in open(src) if l.startswith('"'))]
    hdr = rows[0]
    ki, mi, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    tot = OrderedDict()
    for r in rows[1:]:
        if r[mi] != "gpu__time_duration.sum":
            continue
        ms = float(r[vi].replace(",", "")) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}[r[ui]]
        k = short(r[ki])
        n, t = tot.get(k, (0, 0.0))
        tot[k] = (n + 1, t + ms)
    total = sum(t for _, t in tot.values())
    with open(os.path.join(ROOT, "profiles", f"{tag}_launch_summary.txt"), "w") as f:
        f.write(f"# {header}\n# cold-cache serialised launches under ncu: compare SHARES, not absolutes\n")
        f.write(f"# total {total:.1f} ms over {sum(n for n, _ in tot.values())} launches\n")
        for k, (n, t) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
            f.write(f"{k:<42s} n={n:5d} total={t:10.3f} ms avg={t / n:9.4f} ms share={100 * t / total:5.1f}%\n")
    print(open(os.path.join(ROOT, "profiles", f"{tag}_launch_summary.txt")).read())


def full(tag, header):
    rep = os.path.join(ROOT, "gpurun_out", f"prof_{tag}.ncu-rep")
    if not os.path.exists(rep):
        print("no", rep)
        return
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(l for l in out.splitlines() if l.startswith('"')))
    hdr, units = rows[0], rows[1]
    res = []
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        u = dict(zip(hdr, units))
        e = OrderedDict(kernel=short(d["Kernel Name"]), grid=d.get("Grid Size", ""), block=d.get("Block Size", ""))
        for k in KEEP:
            if k in d:
                e[k] = f"{d[k]} {u[k]}".strip()
        res.append(e)
    json.dump(res, open(os.path.join(ROOT, "profiles", f"{tag}_ncu_full_summary.json"), "w"), indent=1)
    with open(os.path.join(ROOT, "profiles", f"{tag}_ncu_full_summary.txt"), "w") as f:
        f.write(f"# {header}\n# read with: ncu -i prof_{tag}.ncu-rep --page raw --csv   (the .ncu-rep stays in gpurun_out/)\n")
        for e in res:
            f.write(f"\n== {e['kernel']}  grid {e['grid']} block {e['block']}\n")
            for k, v in e.items():
                if k not in ("kernel", "grid", "block"):
                    f.write(f"  {k:<92s} {v}\n")
    print(open(os.path.join(ROOT, "profiles", f"{tag}_ncu_full_summary.txt")).read()[:6000])


if __name__ == "__main__":
    tag = sys.argv[1]
    header = sys.argv[2] if len(sys.argv) > 2 else tag
    launches(tag, header)
    full(tag, header)

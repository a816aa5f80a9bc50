#!/usr/bin/env python
"""Turns the ncu outputs that tools/gpu_profile.sh left in gpurun_out/ into the tracked summaries under profiles/:
  profiles/<tag>_launches.csv          the ncu launch list of `bench.py` (gpu__time_duration per launch) + per-kernel shares
  profiles/<tag>_<kernel>_ncu.txt      key metrics of the --set full capture (time, DRAM bytes, throughputs, stalls, pipes)
  profiles/<tag>_<kernel>_stages.txt   warp-instructions per kernel stage / source line (from the source page)
  profiles/<tag>_k_stencil_march.sass  cuobjdump -sass of the shipped stencil kernel
  profiles/stencil_traffic.json        DRAM bytes per launch of the stencil kernel (bench.py reports it as roofline.traffic)
usage: tools/summarize_profiles.py [tag]"""
import collections
import csv
import io
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TAG = sys.argv[1] if len(sys.argv) > 1 else "r01"
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
os.makedirs(P, exist_ok=True)

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__shared_mem_per_block_dynamic",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tc.avg.pct_of_peak_sustained_active",
        "sm__cycles_active.avg", "gpc__cycles_elapsed.max", "sm__cycles_elapsed.avg.per_second"]


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    return rows[0], rows[1], rows[2:]


def kernel_summary(name, px=None):
    rep = os.path.join(G, f"{TAG}_{name}.ncu-rep")
    if not os.path.exists(rep):
        return None
    hdr, units, vals = raw(rep)
    v = vals[0]
    d = dict(zip(hdr, v))
    u = dict(zip(hdr, units))
    lines = [f"# ncu --set full --clock-control none, one launch of {d.get('Kernel Name')} grid {d.get('Grid Size')} block {d.get('Block Size')}",
             f"# source: gpurun_out/{TAG}_{name}.ncu-rep (command: bench.py --steps 5 --warmup 3, launch skipped 3); times under ncu are cold-cache, serialised"]
    for k in KEYS:
        if k in d:
            lines.append(f"{k:75s} {d[k]:>18s} {u[k]}")
    for k in sorted(d):
        if k.startswith("smsp__average_warps_issue_stalled") and k.endswith("per_issue_active.ratio"):
            lines.append(f"{k:75s} {d[k]:>18s}")
    def num(k):
        return float(d[k].replace(",", ""))
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    traffic = num("dram__bytes_read.sum") * scale[u["dram__bytes_read.sum"]] + num("dram__bytes_write.sum") * scale[u["dram__bytes_write.sum"]]
    tscale = {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1}
    t = num("gpu__time_duration.sum") * tscale[u["gpu__time_duration.sum"]]
    lines.append(f"derived: DRAM traffic {traffic/1e6:.1f} MB per launch, {traffic/t/1e9:.0f} GB/s under ncu")
    inst = num("smsp__inst_executed.sum")
    if px:
        lines.append(f"derived: {inst*32/px:.1f} lane-instructions per pixel ({px} px per launch), algorithmic bytes {3.25*px/1e6:.1f} MB")
    open(os.path.join(P, f"{TAG}_{name}_ncu.txt"), "w").write("\n".join(lines) + "\n")
    # source page -> per-stage split
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
    tmp = f"/tmp/{TAG}_{name}_src.csv"
    open(tmp, "w").write(src)
    if name == "stencil":
        # (by the FUNCTIONS of k_stencil_march.cuh: tools/ncu_march_stages.py keyed stages by line ranges that went stale)
        a = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_by_function.py"), rep, str(px)], capture_output=True, text=True).stdout
    else:
        a = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_stage_split.py"), tmp] + ([str(px)] if px else []), capture_output=True, text=True).stdout
    b = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_source_summary.py"), tmp, "30"], capture_output=True, text=True).stdout
    open(os.path.join(P, f"{TAG}_{name}_stages.txt"), "w").write("# per kernel stage (stencil: source markers; others: code between BAR.SYNC instructions, in SASS address order)\n" + a + "\n# per source line\n" + b)
    return traffic


def launches():
    src = os.path.join(G, f"{TAG}_launches.csv")
    if not os.path.exists(src):
        return
    rows = [r for r in csv.reader(open(src)) if r and r[0].isdigit()]
    per = collections.defaultdict(list)
    for r in rows:
        per[r[4].split("(")[0]].append(float(r[-1]))
    tot = sum(sum(v) for v in per.values())
    with open(os.path.join(P, f"{TAG}_launches.csv"), "w") as f:
        f.write("# ncu --metrics gpu__time_duration.sum --clock-control none python bench.py --steps 5 --warmup 3 --no-cpu\n")
        f.write("# per-kernel share of the GPU time of the whole command (device-path loop + e2e loop):\n")
        for k, v in sorted(per.items(), key=lambda kv: -sum(kv[1])):
            f.write(f"#   {k:40s} launches {len(v):4d}  mean {sum(v)/len(v)/1e3:9.1f} us  share {100*sum(v)/tot:5.1f} %\n")
        f.write(open(src).read())


px = 64 * 1920 * 1080
tr = kernel_summary("stencil", px)
for k in ("k_uf_tile", "k_uf_border", "k_uf_resolve"):
    kernel_summary(k, px)
launches()
if tr:
    json.dump({"bytes_per_launch_batch1080p": tr, "source": f"profiles/{TAG}_stencil_ncu.txt (dram__bytes_read.sum + dram__bytes_write.sum, one launch)"},
              open(os.path.join(P, "stencil_traffic.json"), "w"), indent=1)
so = os.path.join(ROOT, "cudacam_b200", "libb200canny.so")
allsass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
keep, on = [], False
for line in allsass.split("\n"):
    if "Function :" in line:
        on = "k_stencil_march" in line and "ILi3E" in line   # the BGR8 instance
    if on:
        keep.append(line)
open(os.path.join(P, f"{TAG}_k_stencil_march.sass"), "w").write("\n".join(keep) + "\n")
print("profiles written:", sorted(os.listdir(P)))

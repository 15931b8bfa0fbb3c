#!/usr/bin/env python
"""Turn the ncu artefacts a gpurun call brought back (gpurun_out/*.ncu-rep, launches_*.csv) into small text
summaries under profiles/ (tracked).  Usage: python scripts/summarize_ncu.py <tag> [<name>=<rep>...]"""
import collections
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEEP = (
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "sm__inst_executed.sum.per_cycle_elapsed", "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tma.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active",
    "sm__cycles_elapsed.max", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sectors_srcunit_tex_op_read.sum",
    "sm__cycles_elapsed.max.per_second", "sm__icc_request_hit_rate.pct",
)


def raw_metrics(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    if len(rows) < 3:
        return []
    hdr, units = rows[0], rows[1]
    res = []
    for vals in rows[2:]:
        d = {}
        for h, u, v in zip(hdr, units, vals):
            d[h] = (v, u)
        res.append(d)
    return res


def summarize_rep(name, rep, tag):
    kernels = raw_metrics(rep)
    lines = [f"# ncu --set full --clock-control none : {os.path.basename(rep)}", ""]
    traffic = {}
    for d in kernels:
        kname = d.get("Kernel Name", ("?", ""))[0]
        lines.append(f"## {kname}")
        for k in KEEP:
            if k in d:
                lines.append(f"{k:75s} {d[k][0]:>20s} {d[k][1]}")
        stalls = sorted(
            ((float(v[0]), k.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""))
             for k, v in d.items() if k.startswith("smsp__average_warps_issue_stalled_") and v[0] not in ("", "n/a")),
            reverse=True,
        )
        lines.append("top stall reasons (warps per issue-active cycle): " + ", ".join(f"{n}={x:.2f}" for x, n in stalls[:6]))
        try:
            def to_bytes(v, u):
                scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}[u]
                return float(v.replace(",", "")) * scale
            rd = to_bytes(*d["dram__bytes_read.sum"])
            wr = to_bytes(*d["dram__bytes_write.sum"])
            traffic[kname.replace("void ", "").split("(")[0].split("<")[0].split("::")[-1].strip()] = rd + wr
            lines.append(f"dram traffic per launch (read+write): {rd + wr:.4e} bytes")
        except Exception:
            pass
        lines.append("")
    path = os.path.join(ROOT, "profiles", f"{tag}_{name}_ncu_full.txt")
    with open(path, "w") as f:
        f.write("\n".join(lines))
    return traffic


def summarize_launches(csv_path, tag):
    rows = [r for r in csv.reader(open(csv_path)) if len(r) > 10]
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[ui], 1e-3)
        agg.setdefault(r[ki], []).append(v * scale)
    tot = sum(sum(v) for v in agg.values())
    lines = [f"# ncu --metrics gpu__time_duration.sum --clock-control none : {os.path.basename(csv_path)}",
             "# per-launch times are cold-cache and serialised: compare SHARES, not absolutes", "",
             f"{'kernel':90s} {'launches':>8s} {'mean_us':>10s} {'share':>7s}"]
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        lines.append(f"{k[:90]:90s} {len(v):8d} {sum(v) / len(v):10.1f} {sum(v) / tot:7.3f}")
    with open(os.path.join(ROOT, "profiles", f"{tag}_launches.txt"), "w") as f:
        f.write("\n".join(lines) + "\n")


def main():
    tag = sys.argv[1]
    os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
    traffic = {}
    tpath = os.path.join(ROOT, "profiles", "traffic_bytes_per_launch.json")
    if os.path.exists(tpath):
        traffic = json.load(open(tpath))
    for arg in sys.argv[2:]:
        name, rep = arg.split("=")
        if rep.endswith(".csv"):
            summarize_launches(rep, tag)
        else:
            t = summarize_rep(name, rep, tag)
            traffic.update(t)
    traffic["_source"] = f"dram__bytes_read.sum + dram__bytes_write.sum per launch, ncu --set full ({tag})"
    json.dump(traffic, open(tpath, "w"), indent=1)


if __name__ == "__main__":
    main()

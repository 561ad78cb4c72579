#!/usr/bin/env python
"""Digest of this round's ncu exports (run here, no GPU needed):
  profiles/r2_launches_metrics.csv   `ncu --metrics gpu__time_duration.sum,dram__bytes_*,... --csv` of bench.py (second lap)
  profiles/r2_fuse_full_raw.csv      `ncu -i fuse_full.ncu-rep --page raw --csv` of the fuse kernel's --set full capture
-> profiles/r2_fuse_capture.json (read by bench.py for roofline.traffic / issue utilisation) and a text summary."""
import csv
import json
import os
import sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
P = os.path.join(ROOT, "profiles")


def short(name):
    for k in ("fuseKernel", "selectBlocksKernel", "itemCullKernel", "itemCompactKernel", "tileMaxKernel", "tilePyramidKernel"):
        if k in name:
            return k
    return name[:40]


def read_metrics(path):
    rows = list(csv.reader(l for l in open(path) if l.startswith('"')))
    hdr = rows[0]
    i_id, i_name, i_metric, i_unit, i_val = (hdr.index(x) for x in ("ID", "Kernel Name", "Metric Name", "Metric Unit", "Metric Value"))
    launches = defaultdict(dict)
    for r in rows[1:]:
        v = float(r[i_val].replace(",", ""))
        u = r[i_unit]
        if r[i_metric].startswith("gpu__time_duration"):
            v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(u, 1.0)  # -> us
        if r[i_metric].startswith("dram__bytes"):
            v *= {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1.0)
        launches[int(r[i_id])]["name"] = short(r[i_name])
        launches[int(r[i_id])][r[i_metric]] = v
    return [launches[k] for k in sorted(launches)]


def main():
    L = read_metrics(os.path.join(P, "r2_launches_metrics.csv"))
    per = defaultdict(lambda: defaultdict(list))
    for l in L:
        for k, v in l.items():
            if k != "name":
                per[l["name"]][k].append(v)
    n_groups = len(per["fuseKernel"]["gpu__time_duration.sum"])
    lines = ["ncu launch list + metrics of `python bench.py --steps 1 --warmup 1` (hall640, second lap = steady state), %d launches, %d fuse groups" % (len(L), n_groups),
             "per-launch times are cold-cache and serialised under ncu: compare SHARES", "",
             "kernel | launches | mean us | share of serialised time | DRAM read MB | DRAM write MB | warp insts (M) | issue active % | warps active %"]
    tot_t = sum(sum(v["gpu__time_duration.sum"]) for v in per.values())
    group = {"us": 0.0, "dram": 0.0}
    for name, v in sorted(per.items(), key=lambda kv: -sum(kv[1]["gpu__time_duration.sum"])):
        n = len(v["gpu__time_duration.sum"])
        mean = lambda k: sum(v[k]) / len(v[k]) if v.get(k) else float("nan")
        lines.append("%s | %d | %.1f | %.1f %% | %.2f | %.2f | %.2f | %.1f | %.1f" % (
            name, n, mean("gpu__time_duration.sum"), 100 * sum(v["gpu__time_duration.sum"]) / tot_t, mean("dram__bytes_read.sum") / 1e6,
            mean("dram__bytes_write.sum") / 1e6, mean("smsp__inst_executed.sum") / 1e6,
            mean("smsp__issue_active.avg.pct_of_peak_sustained_active"), mean("sm__warps_active.avg.pct_of_peak_sustained_active")))
        per_group = n / max(n_groups, 1)
        group["us"] += mean("gpu__time_duration.sum") * per_group
        group["dram"] += (mean("dram__bytes_read.sum") + mean("dram__bytes_write.sum")) * per_group
    cap = {"frames_per_launch": 32, "source": "profiles/r2_launches_metrics.csv (ncu --metrics, --clock-control none, this round)",
           "dram_bytes_per_launch_group": group["dram"], "serialised_us_per_launch_group": group["us"],
           "fuse_dram_bytes_per_launch": (sum(per["fuseKernel"]["dram__bytes_read.sum"]) + sum(per["fuseKernel"]["dram__bytes_write.sum"])) / max(n_groups, 1),
           "fuse_issue_slot_utilization_pct": sum(per["fuseKernel"]["smsp__issue_active.avg.pct_of_peak_sustained_active"]) / max(n_groups, 1),
           "fuse_warp_instructions_per_launch": sum(per["fuseKernel"]["smsp__inst_executed.sum"]) / max(n_groups, 1)}
    raw = os.path.join(P, "r2_fuse_full_raw.csv")
    if os.path.exists(raw):
        rows = list(csv.reader(open(raw)))
        hdr, units, vals = rows[0], rows[1], rows[2:]
        want = ["smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
                "launch__registers_per_thread", "launch__occupancy_limit_registers", "dram__bytes_read.sum", "dram__bytes_write.sum",
                "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
                "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
                "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio" ]
        lines += ["", "--set full capture of fuseKernel (profiles/r2_fuse_full_raw.csv), mean over %d launches:" % len(vals)]
        stall = {}
        for i, hname in enumerate(hdr):
            try:
                m = sum(float(v[i].replace(",", "")) for v in vals) / len(vals)
            except Exception:
                continue
            if hname in want:
                lines.append("  %s = %.3f %s" % (hname, m, units[i]))
            if hname.startswith("smsp__average_warps_issue_stalled_") and hname.endswith("_per_issue_active.ratio"):
                stall[hname[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]] = m
        if stall:
            tot = sum(stall.values())
            lines.append("  stall reasons (warps per issue-active cycle, share): " + ", ".join("%s %.1f %%" % (k, 100 * v / tot) for k, v in sorted(stall.items(), key=lambda kv: -kv[1])[:8]))
            cap["fuse_long_scoreboard_stall_share_pct"] = 100 * stall.get("long_scoreboard", 0.0) / tot
    json.dump(cap, open(os.path.join(P, "r2_fuse_capture.json"), "w"), indent=1)
    open(os.path.join(P, "r2_ncu_summary.txt"), "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))
    print(json.dumps(cap, indent=1))


if __name__ == "__main__":
    main()

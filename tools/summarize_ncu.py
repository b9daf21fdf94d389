#!/usr/bin/env python
"""Summarise gpurun_out ncu artefacts into profiles/ (tracked).

  python tools/summarize_ncu.py <tag> [--rep gpurun_out/prof_halfstep.ncu-rep] [--launches gpurun_out/launches.csv]

Writes profiles/<tag>_launches.csv (copy of the ncu launch list), profiles/<tag>_summary.json
(per-kernel time shares + the `ncu --set full` key metrics per captured launch).
"""
import argparse
import collections
import csv
import json
import os
import re
import shutil
import subprocess

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
    "launch__block_size", "smsp__inst_executed.sum", "lts__t_sectors_srcunit_tex_op_read.sum",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
]


def to_bytes(v, unit):
    mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    return float(v) * mult.get(unit, 1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("tag")
    ap.add_argument("--rep", default="gpurun_out/prof_halfstep.ncu-rep")
    ap.add_argument("--launches", default="gpurun_out/launches.csv")
    a = ap.parse_args()
    os.makedirs("profiles", exist_ok=True)
    out = {"tag": a.tag}
    if os.path.exists(a.launches):
        shutil.copy(a.launches, f"profiles/{a.tag}_launches.csv")
        rows = list(csv.reader(l for l in open(a.launches) if l.startswith('"')))
        hdr = rows[0]
        iK, iV = hdr.index("Kernel Name"), hdr.index("Metric Value")
        tot, cnt = collections.defaultdict(float), collections.Counter()
        for r in rows[1:]:
            name = re.sub(r"\(.*", "", r[iK])
            tot[name] += float(r[iV].replace(",", "")) / 1e6
            cnt[name] += 1
        out["launch_list_ms"] = {k: {"total_ms": round(v, 4), "launches": cnt[k]} for k, v in
                                 sorted(tot.items(), key=lambda x: -x[1])}
    if os.path.exists(a.rep):
        raw = subprocess.run(["ncu", "-i", a.rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(raw.splitlines()))
        hdr, units, data = rows[0], rows[1], rows[2:]
        caps = []
        for r in data:
            c = {"kernel": r[hdr.index("Kernel Name")]}
            for k in KEYS:
                if k in hdr:
                    i = hdr.index(k)
                    v = r[i].replace(",", "")
                    try:
                        c[k] = to_bytes(v, units[i]) if "bytes" in k else float(v)
                    except ValueError:
                        c[k] = v
                    if "bytes" not in k and units[i]:
                        c[k + ".unit"] = units[i]
            if "dram__bytes_read.sum" in c:
                c["dram_traffic_bytes"] = c["dram__bytes_read.sum"] + c["dram__bytes_write.sum"]
            caps.append(c)
        out["ncu_full_captures"] = caps
    with open(f"profiles/{a.tag}_summary.json", "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out, indent=1)[:3000])


if __name__ == "__main__":
    main()

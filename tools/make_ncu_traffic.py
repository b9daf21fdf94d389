#!/usr/bin/env python
"""profiles/ncu_traffic.json from a tools/summarize_ncu.py digest: DRAM bytes per launch and L2 throughput of the two
solve kernels of the headline configuration (read by bench.py for roofline.traffic / frac_dram / frac_l2).

  python tools/make_ncu_traffic.py profiles/r02o_halfstep_summary.json
"""
import json
import sys

src = sys.argv[1]
d = json.load(open(src))
caps = [c for c in d["ncu_full_captures"] if "half_step_kernel" in c["kernel"]]
h = next(c for c in caps if "tiled" not in c["kernel"])
w = next(c for c in caps if "tiled" in c["kernel"])
short = lambda c: c["kernel"].split("(")[0].replace("void ", "")
out = {
    "1000000x100000x0.001_k64_cholesky_n1": {
        "bytes_per_launch_mean": (h["dram_traffic_bytes"] + w["dram_traffic_bytes"]) / 2.0,
        "H_update_launch": h["dram_traffic_bytes"],
        "W_update_launch": w["dram_traffic_bytes"],
        "H_update_kernel": short(h),
        "W_update_kernel": short(w),
        "lts_throughput_frac_mean": (h["lts__throughput.avg.pct_of_peak_sustained_elapsed"] +
                                     w["lts__throughput.avg.pct_of_peak_sustained_elapsed"]) / 200.0,
        "H_update_lts_throughput_pct": h["lts__throughput.avg.pct_of_peak_sustained_elapsed"],
        "W_update_lts_throughput_pct": w["lts__throughput.avg.pct_of_peak_sustained_elapsed"],
        "H_update_l2_hit_pct": h["lts__t_sector_hit_rate.pct"],
        "W_update_l2_hit_pct": w["lts__t_sector_hit_rate.pct"],
        "H_update_ncu_ms": h["gpu__time_duration.sum"],
        "W_update_ncu_ms": w["gpu__time_duration.sum"],
        "source": f"{src} (ncu --set full --clock-control none; dram__bytes_read.sum + dram__bytes_write.sum, "
                  "lts__throughput.avg.pct_of_peak_sustained_elapsed)",
    }
}
json.dump(out, open("profiles/ncu_traffic.json", "w"), indent=1)
print(json.dumps(out, indent=1))

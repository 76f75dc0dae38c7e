"""Summarises an `ncu --set full` report (read here, on the CPU box) into the small JSON kept under profiles/.

    ncu -i gpurun_out/prof.ncu-rep --page raw --csv > /tmp/raw.csv
    python tools/ncu_summary.py /tmp/raw.csv profiles/r1_ncu_full_<name>.json
"""
import csv
import json
import re
import sys


def main(src, dst):
    rows = list(csv.reader(open(src)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}
    # ncu picks the unit per report (us / ms, Mbyte / Gbyte): normalise to microseconds and megabytes
    SCALE = {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3, "s": 1e6, "second": 1e6,
             "byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}

    def f(r, key, scale=1.0):
        try:
            return float(r[ix[key]]) * scale * SCALE.get(units[ix[key]], 1.0)
        except (KeyError, ValueError):
            return None

    out = []
    for r in data:
        name = r[ix["Kernel Name"]]
        m = re.search(r"(k_\w+)<(.*?)>\(", name)
        short = f"{m.group(1)}<{m.group(2).replace(' ', '').replace('mgb::', '')}>" if m else name
        us = f(r, "gpu__time_duration.sum")
        rd, wr = f(r, "dram__bytes_read.sum"), f(r, "dram__bytes_write.sum")      # MB in the raw page
        inst = f(r, "inst_executed")
        out.append({
            "kernel": short, "grid": int(f(r, "launch__grid_size") or 0), "us": us,
            "dram_read_MB": rd, "dram_write_MB": wr,
            "dram_GBs": round((rd + wr) / us * 1e3, 1) if us and rd is not None else None,
            "dram_pct": f(r, "dram__throughput.avg.pct_of_peak_sustained_elapsed"),
            "lts_pct": f(r, "lts__throughput.avg.pct_of_peak_sustained_elapsed"),
            "l1tex_pct": f(r, "l1tex__throughput.avg.pct_of_peak_sustained_elapsed"),
            "warps_active_pct": f(r, "sm__warps_active.avg.pct_of_peak_sustained_active"),
            "issue_active_pct": f(r, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
            "warp_inst_executed": inst, "regs": int(f(r, "launch__registers_per_thread") or 0),
            "sm_cycles_active": f(r, "sm__cycles_active.avg"),
            "l1_hit_pct": f(r, "l1tex__t_sector_hit_rate.pct"), "l2_hit_pct": f(r, "lts__t_sector_hit_rate.pct"),
            "stall_long_scoreboard": f(r, "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio"),
            "stall_short_scoreboard": f(r, "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio"),
            "stall_barrier": f(r, "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio"),
            "stall_wait": f(r, "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio"),
        })
    json.dump(out, open(dst, "w"), indent=1)
    print(f"{len(out)} launches -> {dst}")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])

#!/usr/bin/env python3
"""Summarise an ncu raw-page CSV (ncu -i X.ncu-rep --page raw --csv) into a small JSON the repo can track.

usage: python profiles/summarize.py gpurun_out/front_raw_v2.csv profiles/r01_front_v2.json
"""
import csv
import json
import sys

KEEP = [
    "Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__cycles_elapsed.max",
]


def main(src, dst):
    rows = list(csv.reader(open(src)))
    hdr, units = rows[0], rows[1]
    out = []
    for r in rows[2:]:
        d = {}
        for k in KEEP:
            if k in hdr:
                i = hdr.index(k)
                d[k] = "%s %s" % (r[i], units[i])
        stalls = {}
        for i, h in enumerate(hdr):
            if "warp_issue_stalled" in h and h.endswith("_per_warp_active.pct"):
                try:
                    v = float(r[i])
                except ValueError:
                    continue
                if v >= 2.0:
                    stalls[h.replace("smsp__warp_issue_stalled_", "").replace("_per_warp_active.pct", "")] = round(v, 2)
        d["stall_pct_of_warp_active(>=2)"] = stalls
        out.append(d)
    json.dump(out, open(dst, "w"), indent=1)
    print("wrote", dst, len(out), "kernels")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])

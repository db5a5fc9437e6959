#!/usr/bin/env python3
"""ncu `--page raw --csv` export -> compact per-kernel summary (the metrics DESIGN.md / profiles/README.md quote + the top stall
reasons), one row per captured launch.  usage: ncu_summary.py <raw.csv> <out.csv>"""
import csv
import sys

KEYS = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__grid_size", "launch__block_size",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static", "launch__occupancy_limit_shared_mem",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_warps", "launch__waves_per_multiprocessor",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "smsp__warps_eligible.avg.per_cycle_active", "smsp__thread_inst_executed_per_inst_executed.ratio"]


def main(src, dst):
    rows = list(csv.reader(open(src)))
    hdr, units = rows[0], rows[1]
    stall = [h for h in hdr if "warps_issue_stalled" in h and h.endswith("_per_issue_active.ratio")]
    out_hdr = [k + (" [%s]" % units[hdr.index(k)] if k in hdr and units[hdr.index(k)] else "") for k in KEYS if k in hdr] + ["top stalls (warps per issue-active cycle)"]
    out = [out_hdr]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        st = sorted(((float(d[h].replace(",", "")), h.split("issue_stalled_")[1].replace("_per_issue_active.ratio", "")) for h in stall if d.get(h) not in (None, "", "n/a")), reverse=True)[:6]
        name = d["Kernel Name"].split("(")[0]
        out.append([name if k == "Kernel Name" else d[k] for k in KEYS if k in hdr] + ["; ".join("%s %.2f" % (n, v) for v, n in st)])
    csv.writer(open(dst, "w", newline="")).writerows(out)


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])

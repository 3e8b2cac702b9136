"""usage: ncu -i file.ncu-rep --page raw --csv | python tools/ncu_summary.py  -> the counters quoted in profiles/*_summary.txt"""
import csv, sys
rows = list(csv.reader(sys.stdin))
hdr = rows[0]
want = ["Kernel Name", "Block Size", "Grid Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_adu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__thread_inst_executed_per_inst_executed.ratio"]
stall = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
units = dict(zip(hdr, rows[1]))
for r in rows[2:]:
    d = dict(zip(hdr, r))
    print("----")
    for w in want:
        if w in d:
            print(f"{w:86s} {d[w]} {units.get(w, '')}")
    for v, s in sorted(((float(d[s] or 0), s) for s in stall), reverse=True)[:8]:
        print(f"{s:86s} {v:.3f}")

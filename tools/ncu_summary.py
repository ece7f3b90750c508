"""Prints selected metrics of an `ncu --page raw --csv` export (one block per profiled launch).
usage: ncu -i rep.ncu-rep --page raw --csv | python tools/ncu_summary.py [extra-metric-substring ...]"""
import csv
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__occupancy_limit",
        "sm__pipe_tensor_cycles_active", "sm__inst_executed_pipe_tensor", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__t_sector_hit_rate.pct",
        "lts__t_sector_hit_rate.pct", "smsp__average_warps_issue_stalled", "sm__pipe_fma_cycles_active.avg.pct", "sm__inst_executed_pipe_xu.sum",
        "smsp__inst_executed_pipe_lsu.sum", "launch__shared_mem_per_block", "sm__maximum_warps_per_active_cycle_pct", "launch__waves_per_multiprocessor"]
rows = list(csv.reader(sys.stdin))
hdr, units = rows[0], rows[1]
want = WANT + sys.argv[1:]
ki = hdr.index("Kernel Name")
for r in rows[2:]:
    print("==", r[ki][:90])
    for i, h in enumerate(hdr):
        name = h.split(".", 2)[-1] if h.count(".") >= 2 and h.split(".")[1].startswith("Triage") else h
        if any(w in h for w in want) and "Triage" not in h:
            print(f"  {h:95s} {r[i]:>16s} {units[i]}")

"""Print selected metrics from an `ncu --page raw --csv` export (one column per metric, one row per launch)."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
want = ["Kernel Name", "gpu__time_duration.sum", "sm__cycles_elapsed.max", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "sm__inst_executed_pipe_tc", 
        "sm__pipe_tc_cycles_active", "sm__pipe_tensor", "sm__inst_executed.avg.per_cycle_elapsed", "sm__inst_issued",
        "smsp__issue_active.avg.pct", "sm__pipe_fma_cycles_active", "sm__pipe_alu_cycles_active", "sm__pipe_xu_cycles_active",
        "smsp__inst_executed_pipe_xu", "sm__throughput.avg.pct", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "sm__pipe_fmaheavy", "sm__pipe_fmalite", "smsp__warp_issue_stalled", "launch__registers_per_thread", "sm__warps_active.avg.pct",
        "lts__t_sector_hit_rate", "sm__cycles_active.avg", "smsp__cycles_active.avg", "sm__pipe_shared_cycles_active", "pipe_tmem", "uniform"]
extra = sys.argv[2:]
for r in rows[2:]:
    print("=" * 100)
    for i, h in enumerate(hdr):
        if any(w in h for w in want + extra):
            v = r[i]
            if h == "Kernel Name": v = v[:90]
            print(f"{h:80s} {units[i]:12s} {v}")

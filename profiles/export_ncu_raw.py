"""Turn `ncu -i <rep> --page raw --csv` into the compact per-launch `metric,value,unit` listing committed under profiles/.
usage: ncu -i gpurun_out/x.ncu-rep --page raw --csv | python profiles/export_ncu_raw.py "<header comment>" [kernel-substring] > profiles/x_ncu_raw.csv"""
import csv, sys
KEEP = ("dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__time_duration.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "launch__shared_mem_per_block_static", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "launch__waves_per_multiprocessor", "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__cycles_elapsed.max", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__ops_path_tensor_src_fp64.avg.pct_of_peak_sustained_elapsed", "sm__ops_path_tensor_src_fp64.min", "sm__ops_path_tensor_src_fp64.max",
        "sm__ops_path_tensor_src_fp64.avg", "sm__cycles_active.avg")
rows = list(csv.reader(sys.stdin))
hdr, units = rows[0], rows[1]
ki = hdr.index("Kernel Name")
print("# " + sys.argv[1])
sub = sys.argv[2] if len(sys.argv) > 2 else ""
for row in rows[2:]:
    if sub and sub not in row[ki]:
        continue
    print(f"Kernel Name,{row[ki]},")
    for i, h in enumerate(hdr):
        if h in KEEP or h.startswith("smsp__average_warps_issue_stalled_"):
            print(f"{h},{row[i]},{units[i]}")
    print()

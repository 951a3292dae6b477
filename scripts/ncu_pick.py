"""filter for `ncu --page raw --csv`: keeps the columns the roofline arithmetic needs (DRAM bytes, durations, pipe utilisation)"""
import csv, sys
keep = ("Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__grid_size", "launch__block_size", "lts__t_bytes.sum")
rows = list(csv.reader(sys.stdin))
hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
cols = [i for i, c in enumerate(rows[hdr]) if c in keep]
w = csv.writer(sys.stdout)
for r in rows[hdr:]:
    if len(r) > max(cols):
        w.writerow([r[i] for i in cols])

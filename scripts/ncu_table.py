"""One row per captured kernel from an `ncu --page raw --csv` dump: the metrics the roofline discussion cites."""
import csv, sys
KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "sm__cycles_elapsed.max", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed"]
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
idx = [hdr.index(k) for k in KEYS if k in hdr]
ki = hdr.index("Kernel Name")
print("kernel".ljust(44) + " | ".join("%s [%s]" % (hdr[i].split(".avg")[0], units[i]) for i in idx))
for r in rows[2:]:
    print(r[ki][:43].ljust(44) + " | ".join(r[i] for i in idx))

"""Summarise an .ncu-rep: key raw metrics + the hottest source lines by sampled stalls."""
import csv, subprocess, sys, collections, io
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, vals = rows[0], rows[-1]
keys = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__occupancy_limit",
        "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "sm__cycles_elapsed.max",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__pcsamp_warps_issue_stalled"]
for h, v in zip(hdr, vals):
    if any(h.startswith(k) or k == h for k in keys) and "not_issued" not in h:
        print("%-90s %s" % (h, v))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
if len(rows) > 2:
    hdr = rows[0]
    try:
        si = hdr.index("# Samples") if "# Samples" in hdr else [i for i, h in enumerate(hdr) if "Samples" in h][0]
    except Exception:
        si = None
    srci = 1 if len(hdr) > 1 else 0
    print("source columns:", hdr[:12])
    if si is not None:
        tot = 0
        items = []
        for r in rows[1:]:
            try:
                n = float(r[si])
            except Exception:
                continue
            tot += n
            items.append((n, r[0], r[srci][:110]))
        items.sort(reverse=True)
        print("total samples", tot)
        for n, a, t in items[:int(sys.argv[2]) if len(sys.argv) > 2 else 25]:
            print("%7.0f %5.1f%%  %s | %s" % (n, 100 * n / max(tot, 1), a, t))

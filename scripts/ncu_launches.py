"""Per-kernel totals of an ncu `--metrics gpu__time_duration.sum --csv` launch list (share of the profiled run)."""
import csv, sys, collections
rows = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if not l.startswith("==")]
rd = csv.reader(lines)
hdr = next(rd)
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
tot = collections.defaultdict(lambda: [0, 0.0])
for r in rd:
    if len(r) <= vi:
        continue
    v = float(r[vi].replace(",", ""))
    v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[ui], 1e-6)
    name = r[ki].split("(")[0][:70]
    tot[name][0] += 1
    tot[name][1] += v
all_ms = sum(v[1] for v in tot.values())
print("kernels %d  launches %d  total %.1f ms (cold-cache, serialised: compare SHARES, not absolutes)" % (len(tot), sum(v[0] for v in tot.values()), all_ms))
print("%-72s %8s %10s %8s %7s" % ("kernel", "launches", "ms", "avg_us", "share"))
for k, (n, ms) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print("%-72s %8d %10.2f %8.1f %6.2f%%" % (k, n, ms, 1e3 * ms / n, 100 * ms / all_ms))

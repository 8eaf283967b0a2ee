"""Hottest SASS instructions of an .ncu-rep by sampled warp stalls (source page)."""
import csv, subprocess, sys, io
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]
si = hdr.index("# Samples"); ei = hdr.index("Instructions Executed")
items = []; tot = 0
for idx, r in enumerate(rows[2:]):
    try: n = float(r[si])
    except Exception: continue
    tot += n
    items.append((n, idx, r[1].strip()[:100], r[ei]))
print("total samples", tot, "instructions", len(items))
for n, idx, t, e in sorted(items, reverse=True)[:top]:
    print("%7.0f %5.1f%%  #%-5d exec %-9s %s" % (n, 100 * n / max(tot, 1), idx, e, t))

"""Per CUDA-source-line stall samples of an .ncu-rep (needs -lineinfo + --import-source on).
usage: ncu_lines.py rep [top] [kernel-substring]"""
import csv, subprocess, sys, io
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
cur_file = None; hdr = None; items = []; cur_fn = ''
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur_file = r[1].split("/")[-1]; continue
    if r[0] == "Function Name": cur_fn = r[1].split("(")[0][-40:]; continue
    if r[0] == "Line No": hdr = r; si = hdr.index("# Samples"); continue
    if hdr is None or r[0] in ("", "..."): continue
    try: n = float(r[si])
    except Exception: continue
    stalls = {}
    for h, v in zip(hdr, r):
        if h.startswith("stall_") and "Not Issued" not in h:
            try:
                if float(v) > 0: stalls[h[6:]] = float(v)
            except Exception: pass
    items.append((n, cur_fn + ' ' + str(cur_file), r[0], r[1].strip()[:95], stalls))
tot = sum(i[0] for i in items)
print("total samples", tot)
for n, f, ln, text, st in sorted(items, key=lambda x: -x[0])[:top]:
    s = " ".join("%s=%d" % (k, v) for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:3])
    print("%7.0f %5.1f%% %s:%s | %s | %s" % (n, 100 * n / max(tot, 1), f, ln, text, s))

#!/bin/bash
OUT=gpurun_out/r3e; mkdir -p $OUT
timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:attn_fwd3|attn_bwd2' -f -o $OUT/hot_attn python scripts/prof_target.py attn > $OUT/prof_attn.log 2>&1; tail -2 $OUT/prof_attn.log
ncu -i $OUT/hot_attn.ncu-rep --page raw --csv > $OUT/hot_attn_raw.csv 2>/dev/null
python scripts/ncu_table.py $OUT/hot_attn_raw.csv | cut -c1-260
python scripts/ncu_lines.py $OUT/hot_attn.ncu-rep 40 > $OUT/hot_attn_lines.txt 2>&1
python - <<'PY'
import csv
rows = list(csv.reader(open("gpurun_out/r3e/hot_attn_raw.csv")))
hdr = rows[0]
for r in rows[2:]:
    print(r[hdr.index("Kernel Name")][:30])
    for i, h in enumerate(hdr):
        if ("pipe" in h and "pct" in h) or "issue_active" in h or "warp_issue_stalled" in h and "pct" in h:
            try:
                if float(r[i]) > 3: print("   %-95s %s" % (h, r[i]))
            except Exception: pass
PY

#!/bin/bash
# what the driver runs at round end (GPU test suite, smoke(), default bench line, reference arm) + the ncu evidence of the final build
OUT=gpurun_out/r2final2; mkdir -p $OUT
timeout 1500 python -m pytest tests -x -q -m gpu -p no:cacheprovider > $OUT/pytest_gpu.log 2>&1; tail -6 $OUT/pytest_gpu.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; tail -3 $OUT/smoke.log
timeout 900 python bench.py --shapes $OUT/shapes.txt > $OUT/bench.json 2> $OUT/bench.err; tail -c 500 $OUT/bench.err; python scripts/bench_brief.py $OUT/bench.json | head -14
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/bench_ref.json 2> $OUT/bench_ref.err; tail -c 600 $OUT/bench_ref.json
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2final2/bench.json"))
print({k: d.get(k) for k in ("value", "ms_per_step", "e2e", "gpu_launches", "launch_mode", "parity_rel_err", "cpu_baseline", "clocks")})
print(d["roofline"])
PY
timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:attn_fwd3|attn_bwd2|attn_fwd_kernel|attn_bwd_kernel' -f -o $OUT/hot_attn python scripts/prof_target.py attn > $OUT/prof_attn.log 2>&1; tail -1 $OUT/prof_attn.log
ncu -i $OUT/hot_attn.ncu-rep --page raw --csv > $OUT/hot_attn_raw.csv 2>/dev/null
python scripts/ncu_table.py $OUT/hot_attn_raw.csv > $OUT/hot_attn_table.txt; cut -c1-230 $OUT/hot_attn_table.txt
python scripts/ncu_lines.py $OUT/hot_attn.ncu-rep 50 > $OUT/hot_attn_lines.txt 2>&1; rm -f $OUT/hot_attn.ncu-rep
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/launches.csv python bench.py --batch 16 --steps 1 --warmup 0 --no-e2e --no-cpu-baseline --no-roofline --graph 0 > $OUT/list_bench.log 2>&1
python scripts/ncu_launches.py $OUT/launches.csv > $OUT/launches_summary.txt 2>&1; rm -f $OUT/launches.csv; head -30 $OUT/launches_summary.txt

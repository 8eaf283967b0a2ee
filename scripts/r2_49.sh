#!/bin/bash
OUT=gpurun_out/r3z; mkdir -p $OUT
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/launches.csv python bench.py --batch 16 --steps 1 --warmup 0 --no-e2e --no-cpu-baseline --no-roofline --graph 0 > $OUT/list_bench.log 2>&1
python scripts/ncu_launches.py $OUT/launches.csv > $OUT/launches_summary.txt 2>&1; rm -f $OUT/launches.csv; head -12 $OUT/launches_summary.txt

#!/bin/bash
OUT=gpurun_out/r3k; mkdir -p $OUT
for mb in 0 40 80 0 40; do echo "== SIDLSG_GN_L2MB=$mb"; SIDLSG_GN_L2MB=$mb timeout 600 python bench.py --no-cpu-baseline --no-e2e --no-roofline --steps 4 > $OUT/bench_$mb.json 2> $OUT/bench_$mb.err; python scripts/bench_brief.py $OUT/bench_$mb.json | head -1; done

#!/bin/bash
# BASELINE configs 4 and 5 with the final build
OUT=gpurun_out/r3y; mkdir -p $OUT
timeout 400 python bench.py --no-cpu-baseline --no-roofline --steps 3 --model SD21_BASE --kappa 2.0 > $OUT/bench_sd21.json 2> $OUT/bench_sd21.err; tail -c 200 $OUT/bench_sd21.err; python scripts/bench_brief.py $OUT/bench_sd21.json | head -2
timeout 400 python bench.py --no-cpu-baseline --no-roofline --steps 3 --num-steps 4 --batch 16 --batch-gpu 16 > $OUT/bench_4step.json 2> $OUT/bench_4step.err; tail -c 200 $OUT/bench_4step.err; python scripts/bench_brief.py $OUT/bench_4step.json | head -2

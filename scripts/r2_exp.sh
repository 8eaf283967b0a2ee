#!/bin/bash
OUT=gpurun_out/r2l; mkdir -p $OUT
for kb in 5 10; do echo "== SIDLSG_BM2_MINKB=$kb"; SIDLSG_BM2_MINKB=$kb timeout 600 python bench.py --no-cpu-baseline --no-e2e --steps 3 --shapes $OUT/shapes_kb$kb.txt > $OUT/bench_kb$kb.json 2> $OUT/bench_kb$kb.err; python scripts/bench_brief.py $OUT/bench_kb$kb.json | head -5; done

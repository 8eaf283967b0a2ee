#!/bin/bash
OUT=gpurun_out/r2x; mkdir -p $OUT
timeout 900 python -m pytest tests/test_tc_gpu.py tests/test_unet_step_gpu.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -2
for f4 in 0 1; do SIDLSG_ATTN_FWD4=$f4 timeout 900 python bench.py --no-cpu-baseline --no-e2e --steps 3 --shapes $OUT/shapes_f4$f4.txt > $OUT/bench_f4$f4.json 2> $OUT/bench_f4$f4.err; tail -c 300 $OUT/bench_f4$f4.err; python scripts/bench_brief.py $OUT/bench_f4$f4.json | head -9; done

#!/bin/bash
OUT=gpurun_out/r2z; mkdir -p $OUT
timeout 900 python -m pytest tests/test_unet_step_gpu.py tests/test_ops_gpu.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -2
timeout 900 python bench.py --no-cpu-baseline --steps 4 --shapes $OUT/shapes.txt > $OUT/bench.json 2> $OUT/bench.err; tail -c 300 $OUT/bench.err; python scripts/bench_brief.py $OUT/bench.json | head -12

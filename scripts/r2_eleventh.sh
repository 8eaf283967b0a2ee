#!/bin/bash
OUT=gpurun_out/r2o; mkdir -p $OUT
timeout 900 python -m pytest tests/test_unet_step_gpu.py tests/test_ops_gpu.py -m gpu -q -x -p no:cacheprovider > $OUT/pytest.log 2>&1; tail -8 $OUT/pytest.log
timeout 900 python bench.py --no-cpu-baseline --steps 5 --shapes $OUT/shapes.txt > $OUT/bench.json 2> $OUT/bench.err; tail -c 600 $OUT/bench.err; python scripts/bench_brief.py $OUT/bench.json | head -3
timeout 900 python bench.py --no-cpu-baseline --steps 5 --graph 0 --no-roofline > $OUT/bench_eager.json 2> $OUT/bench_eager.err; python scripts/bench_brief.py $OUT/bench_eager.json | head -2

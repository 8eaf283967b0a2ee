#!/bin/bash
OUT=gpurun_out/r2k; mkdir -p $OUT
timeout 900 python -m pytest tests/test_ops_gpu.py tests/test_tc_gpu.py tests/test_unet_step_gpu.py -m gpu -q -x -p no:cacheprovider > $OUT/pytest.log 2>&1; tail -12 $OUT/pytest.log
timeout 300 python scripts/micro.py gn 10 2>&1 | tee $OUT/micro_gn.txt
timeout 300 python scripts/micro.py attn 10 2>&1 | head -2 | tee $OUT/micro_attn.txt
timeout 900 python bench.py --no-cpu-baseline --steps 3 --shapes $OUT/shapes.txt > $OUT/bench.json 2> $OUT/bench.err; tail -c 400 $OUT/bench.err; python scripts/bench_brief.py $OUT/bench.json

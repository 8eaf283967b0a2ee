#!/bin/bash
OUT=gpurun_out/r2i; mkdir -p $OUT
timeout 900 python -m pytest tests/test_tc_gpu.py tests/test_unet_step_gpu.py -m gpu -q -x -p no:cacheprovider > $OUT/pytest.log 2>&1; tail -8 $OUT/pytest.log
timeout 300 python scripts/micro.py attn 10 2>&1 | tee $OUT/micro_attn.txt
timeout 120 python scripts/trace_attn_fwd.py > $OUT/trace_fwd3.txt 2>&1; tail -5 $OUT/trace_fwd3.txt
timeout 120 python scripts/trace_attn_bwd.py > $OUT/trace_bwd2.txt 2>&1; tail -4 $OUT/trace_bwd2.txt
timeout 900 python bench.py --no-cpu-baseline --steps 3 --shapes $OUT/shapes.txt > $OUT/bench.json 2> $OUT/bench.err; tail -c 400 $OUT/bench.err; python scripts/bench_brief.py $OUT/bench.json

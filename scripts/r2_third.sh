#!/bin/bash
OUT=gpurun_out/r2c; mkdir -p $OUT
timeout 900 python -m pytest tests/test_tc_gpu.py tests/test_training_loop_gpu.py tests/test_ops_gpu.py -m gpu -q -x -p no:cacheprovider > $OUT/pytest.log 2>&1; tail -8 $OUT/pytest.log
for f in 0 1; do echo "== SIDLSG_ATTN_FWD3=$f"; SIDLSG_ATTN_FWD3=$f timeout 300 python scripts/micro.py attn 10 2>&1 | tee $OUT/micro_attn_fwd3_$f.txt; done
timeout 900 python bench.py --no-cpu-baseline --steps 3 --shapes $OUT/shapes.txt > $OUT/bench.json 2> $OUT/bench.err; tail -c 400 $OUT/bench.err; python scripts/bench_brief.py $OUT/bench.json
timeout 900 python bench.py --no-cpu-baseline --steps 3 --model SD21_BASE --kappa 2.0 --shapes $OUT/shapes_sd21.txt > $OUT/bench_sd21.json 2> $OUT/bench_sd21.err; tail -c 400 $OUT/bench_sd21.err; python scripts/bench_brief.py $OUT/bench_sd21.json
timeout 900 python bench.py --no-cpu-baseline --steps 3 --num-steps 4 --batch 16 --batch-gpu 16 --shapes $OUT/shapes_4step.txt > $OUT/bench_4step.json 2> $OUT/bench_4step.err; tail -c 400 $OUT/bench_4step.err; python scripts/bench_brief.py $OUT/bench_4step.json

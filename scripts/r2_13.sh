#!/bin/bash
# forward attention with eight softmax warps per query tile (attn_fwd4_kernel) against the four-warp kernel
OUT=gpurun_out/r2p; mkdir -p $OUT
timeout 900 python -m pytest tests/test_tc_gpu.py -m gpu -q -x -p no:cacheprovider > $OUT/pytest.log 2>&1; tail -5 $OUT/pytest.log
for f4 in 0 1; do echo "== SIDLSG_ATTN_FWD4=$f4"; SIDLSG_ATTN_FWD4=$f4 timeout 300 python scripts/micro.py attn 10 2>&1 | grep -A1 fwd | tee $OUT/micro_fwd4_$f4.txt; done
timeout 900 python bench.py --no-cpu-baseline --steps 3 --shapes $OUT/shapes.txt > $OUT/bench.json 2> $OUT/bench.err; tail -c 400 $OUT/bench.err; python scripts/bench_brief.py $OUT/bench.json | head -12

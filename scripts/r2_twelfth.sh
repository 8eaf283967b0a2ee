#!/bin/bash
OUT=gpurun_out/r2p; mkdir -p $OUT
timeout 900 python -m pytest tests/test_tc_gpu.py tests/test_ops_gpu.py tests/test_unet_step_gpu.py -m gpu -q -x -p no:cacheprovider > $OUT/pytest.log 2>&1; tail -8 $OUT/pytest.log
for f in 0 1; do echo "== SIDLSG_BRES=$f"; SIDLSG_BRES=$f timeout 300 python scripts/micro.py gemm 10 2>&1 | grep "fwd\|dgrad" | head -8 | tee $OUT/micro_gemm_$f.txt; done
timeout 300 python scripts/micro.py gn 10 2>&1 | tee $OUT/micro_gn.txt
timeout 900 python bench.py --no-cpu-baseline --steps 5 --shapes $OUT/shapes.txt > $OUT/bench.json 2> $OUT/bench.err; tail -c 600 $OUT/bench.err; python scripts/bench_brief.py $OUT/bench.json | head -12

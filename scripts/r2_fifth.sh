#!/bin/bash
OUT=gpurun_out/r2e; mkdir -p $OUT
timeout 900 python -m pytest tests/test_tc_gpu.py tests/test_ops_gpu.py tests/test_unet_step_gpu.py -m gpu -q -x -p no:cacheprovider > $OUT/pytest.log 2>&1; tail -8 $OUT/pytest.log
for pz in 0 4 6 8; do echo "== SIDLSG_ATTN_POLY=$pz"; SIDLSG_ATTN_POLY=$pz timeout 300 python scripts/micro.py attn 10 2>&1 | grep fwd | tee -a $OUT/micro_poly_$pz.txt; done
for b2 in 0 1; do echo "== SIDLSG_ATTN_BWD2=$b2"; SIDLSG_ATTN_BWD2=$b2 timeout 300 python scripts/micro.py attn 10 2>&1 | grep -A1 fwd | tee -a $OUT/micro_bwd2_$b2.txt; done
timeout 120 python scripts/trace_attn_fwd.py > $OUT/trace_fwd3_poly6.txt 2>&1; tail -4 $OUT/trace_fwd3_poly6.txt
timeout 900 python bench.py --no-cpu-baseline --steps 3 --shapes $OUT/shapes.txt > $OUT/bench.json 2> $OUT/bench.err; tail -c 400 $OUT/bench.err; python scripts/bench_brief.py $OUT/bench.json

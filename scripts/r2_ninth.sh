#!/bin/bash
OUT=gpurun_out/r2j; mkdir -p $OUT
timeout 900 python -m pytest tests/test_tc_gpu.py -m gpu -q -x -p no:cacheprovider > $OUT/pytest.log 2>&1; tail -4 $OUT/pytest.log
timeout 300 python scripts/micro.py attn 10 2>&1 | tee $OUT/micro_attn.txt
timeout 120 python scripts/trace_attn_fwd.py > $OUT/trace_fwd3.txt 2>&1; tail -4 $OUT/trace_fwd3.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/launches.csv python bench.py --batch 16 --steps 1 --warmup 0 --no-e2e --no-cpu-baseline --no-roofline > $OUT/list_bench.log 2>&1
python scripts/ncu_launches.py $OUT/launches.csv > $OUT/launches_summary.txt 2>&1; rm -f $OUT/launches.csv; head -60 $OUT/launches_summary.txt

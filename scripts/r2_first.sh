#!/bin/bash
# round-2 first GPU visit: elect.sync validation (parity + speed) against the lane-0 build
OUT=gpurun_out/r2a; mkdir -p $OUT
timeout 120 ./sid_lsg_b200/_C/ubench_mma > $OUT/ubench_mma.txt 2>&1
SIDLSG_TEST_ELECT=1 timeout 900 python -m pytest tests/test_tc_gpu.py -m gpu -q -p no:cacheprovider > $OUT/pytest_tc_default.log 2>&1; tail -3 $OUT/pytest_tc_default.log
SIDLSG_ELECT=1 timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider > $OUT/pytest_elect.log 2>&1; tail -3 $OUT/pytest_elect.log
for e in 0 1; do
  SIDLSG_ELECT=$e timeout 600 python scripts/micro.py all 5 > $OUT/micro_elect$e.txt 2>&1
  SIDLSG_ELECT=$e timeout 900 python bench.py --no-cpu-baseline --steps 3 --shapes $OUT/shapes_elect$e.txt > $OUT/bench_elect$e.json 2> $OUT/bench_elect$e.err
  python scripts/bench_brief.py $OUT/bench_elect$e.json
done

#!/bin/bash
# round-2 second GPU visit: whole GPU suite (new surface / split3 / full-size parity tests), then bench lines
OUT=gpurun_out/r2b; mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q -x -p no:cacheprovider --deselect tests/test_fullsize_gpu.py > $OUT/pytest.log 2>&1; tail -15 $OUT/pytest.log
timeout 1500 python -m pytest tests/test_fullsize_gpu.py -m gpu -q -s -p no:cacheprovider > $OUT/pytest_fullsize.log 2>&1; grep -E "parity|passed|failed|Error|error" $OUT/pytest_fullsize.log | tail -20
timeout 600 python __graft_entry__.py smoke > $OUT/smoke.log 2>&1; tail -3 $OUT/smoke.log
timeout 900 python bench.py --steps 3 --shapes $OUT/shapes.txt > $OUT/bench.json 2> $OUT/bench.err; tail -c 600 $OUT/bench.err; python scripts/bench_brief.py $OUT/bench.json

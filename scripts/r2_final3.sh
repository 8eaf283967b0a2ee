#!/bin/bash
# what the driver runs at round end, on the final build: GPU test suite, smoke(), default bench line, reference arm
OUT=gpurun_out/r2final3; mkdir -p $OUT
timeout 1500 python -m pytest tests -x -q -m gpu -p no:cacheprovider > $OUT/pytest_gpu.log 2>&1; tail -4 $OUT/pytest_gpu.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; tail -3 $OUT/smoke.log
timeout 900 python bench.py --shapes $OUT/shapes.txt > $OUT/bench.json 2> $OUT/bench.err; tail -c 300 $OUT/bench.err; python scripts/bench_brief.py $OUT/bench.json | head -14
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/bench_ref.json 2> $OUT/bench_ref.err; cut -c1-300 $OUT/bench_ref.json

#!/bin/bash
# what the driver runs at round end: GPU test suite, smoke(), the default bench line, the reference arm
OUT=gpurun_out/r2final; mkdir -p $OUT
timeout 1500 python -m pytest tests -x -q -m gpu -p no:cacheprovider > $OUT/pytest_gpu.log 2>&1; tail -6 $OUT/pytest_gpu.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; tail -3 $OUT/smoke.log
timeout 900 python bench.py --shapes $OUT/shapes.txt > $OUT/bench.json 2> $OUT/bench.err; tail -c 500 $OUT/bench.err; python scripts/bench_brief.py $OUT/bench.json | head -14
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2final/bench.json"))
print({k: d.get(k) for k in ("value", "ms_per_step", "e2e", "gpu_launches", "launch_mode", "parity_rel_err", "cpu_baseline", "clocks")})
print(d["roofline"])
PY

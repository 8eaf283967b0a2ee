#!/bin/bash
OUT=gpurun_out/r3x; mkdir -p $OUT
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; tail -2 $OUT/smoke.log
timeout 600 python -m pytest tests/test_tc_gpu.py tests/test_unet_step_gpu.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -2

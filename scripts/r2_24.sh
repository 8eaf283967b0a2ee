#!/bin/bash
# forward attention: S buffer handed back before the last quarter of the exponentials
OUT=gpurun_out/r3a; mkdir -p $OUT
timeout 600 python -m pytest tests/test_tc_gpu.py -m gpu -q -x -p no:cacheprovider -k attention 2>&1 | tail -2
timeout 300 python scripts/micro.py attn 10 2>&1 | grep -A1 "attn fwd" | tee $OUT/micro.txt
timeout 120 python scripts/trace_attn_fwd.py > $OUT/trace_fwd3.txt 2>&1; cut -c1-200 $OUT/trace_fwd3.txt | tail -9

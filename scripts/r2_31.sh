#!/bin/bash
# attention backward with two MMA issuer warps
OUT=gpurun_out/r3h; mkdir -p $OUT
timeout 600 python -m pytest tests/test_tc_gpu.py -m gpu -q -x -p no:cacheprovider -k attention 2>&1 | tail -3
timeout 300 python scripts/micro.py attn 10 2>&1 | grep -A1 "attn fwd" | tee $OUT/micro.txt
timeout 120 python scripts/trace_attn_bwd.py > $OUT/trace_bwd2.txt 2>&1; cut -c1-200 $OUT/trace_bwd2.txt | tail -14

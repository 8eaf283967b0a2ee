#!/bin/bash
# attention after the code-size diet (wait loops not unrolled, one copy of the exponential code, compact MMA loop)
OUT=gpurun_out/r2s; mkdir -p $OUT
timeout 600 python -m pytest tests/test_tc_gpu.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -2
for f4 in 0 1; do echo "== SIDLSG_ATTN_FWD4=$f4"; SIDLSG_ATTN_FWD4=$f4 timeout 300 python scripts/micro.py attn 10 2>&1 | grep -A1 "attn fwd" | tee $OUT/micro_fwd4_$f4.txt; done
timeout 300 python scripts/micro.py gemm 10 2>&1 | tee $OUT/micro_gemm.txt | tail -30
SIDLSG_ATTN_FWD4=0 timeout 120 python scripts/trace_attn_fwd.py > $OUT/trace_fwd3.txt 2>&1; cut -c1-200 $OUT/trace_fwd3.txt | tail -9

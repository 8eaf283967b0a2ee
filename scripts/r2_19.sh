#!/bin/bash
OUT=gpurun_out/r2v; mkdir -p $OUT
timeout 600 python -m pytest tests/test_tc_gpu.py -m gpu -q -x -p no:cacheprovider -k attention 2>&1 | tail -2
for f4 in 0 1; do echo "== FWD4=$f4"; SIDLSG_ATTN_FWD4=$f4 timeout 300 python scripts/micro.py attn 10 2>&1 | grep "attn fwd" | tee $OUT/micro_fwd4_$f4.txt; done

#!/bin/bash
# forward attention: helper warps on hardware warps 0/1, suspend hint on the softmax warps' score wait
OUT=gpurun_out/r2r; mkdir -p $OUT
export SIDLSG_ATTN_FWD4=0
for hf in 0 1; do for sh in 0 20 100; do echo "== HFIRST=$hf SOFT_HINT=$sh"; SIDLSG_ATTN_HFIRST=$hf SIDLSG_ATTN_SOFT_HINT=$sh timeout 300 python scripts/micro.py attn 10 2>&1 | grep "attn fwd" | tee $OUT/micro_hf${hf}_sh$sh.txt; done; done
SIDLSG_ATTN_HFIRST=1 timeout 300 python -m pytest tests/test_tc_gpu.py -m gpu -q -x -p no:cacheprovider -k attention 2>&1 | tail -2
SIDLSG_ATTN_HFIRST=1 SIDLSG_ATTN_SOFT_HINT=20 timeout 120 python scripts/trace_attn_fwd.py > $OUT/trace_hf1_sh20.txt 2>&1; cut -c1-200 $OUT/trace_hf1_sh20.txt | tail -9

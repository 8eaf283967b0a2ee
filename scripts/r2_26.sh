#!/bin/bash
OUT=gpurun_out/r3c; mkdir -p $OUT
for sw in 0 1; do echo "== swap $sw"; SIDLSG_ATTN_SWAP=$sw timeout 120 python scripts/trace_attn_fwd.py > $OUT/trace_swap$sw.txt 2>&1; cut -c1-220 $OUT/trace_swap$sw.txt | tail -6; SIDLSG_ATTN_SWAP=$sw timeout 300 python scripts/micro.py attn 10 2>&1 | grep "attn fwd" | head -1; done

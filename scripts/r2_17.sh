#!/bin/bash
OUT=gpurun_out/r2t; mkdir -p $OUT
SIDLSG_ATTN_FWD4=0 timeout 120 python scripts/trace_attn_fwd.py > $OUT/trace_fwd3_detail.txt 2>&1; cut -c1-220 $OUT/trace_fwd3_detail.txt | tail -26

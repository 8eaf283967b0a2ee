#!/bin/bash
OUT=gpurun_out/r3g; mkdir -p $OUT
timeout 120 python scripts/trace_attn_bwd.py > $OUT/trace_bwd2.txt 2>&1; cut -c1-200 $OUT/trace_bwd2.txt | tail -24

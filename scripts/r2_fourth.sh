#!/bin/bash
OUT=gpurun_out/r2d; mkdir -p $OUT
timeout 120 ./sid_lsg_b200/_C/ubench_mma m > $OUT/ubench_mma_major.txt 2>&1; cat $OUT/ubench_mma_major.txt
timeout 120 python scripts/trace_attn_fwd.py > $OUT/trace_fwd3.txt 2>&1; cat $OUT/trace_fwd3.txt
timeout 120 python scripts/trace_attn_bwd.py > $OUT/trace_bwd.txt 2>&1; cat $OUT/trace_bwd.txt

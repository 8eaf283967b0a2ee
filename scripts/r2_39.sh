#!/bin/bash
OUT=gpurun_out/r3p; mkdir -p $OUT
timeout 120 python scripts/trace_gemm.py 262144 320 320 > $OUT/trace_gemm_320.txt 2>&1; cut -c1-200 $OUT/trace_gemm_320.txt | tail -8
timeout 300 python scripts/micro.py gemm 20 2>&1 | grep "linear fwd" | tee $OUT/micro_gemm.txt

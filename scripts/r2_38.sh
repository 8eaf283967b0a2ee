#!/bin/bash
OUT=gpurun_out/r3o; mkdir -p $OUT
timeout 120 python scripts/trace_gemm.py 262144 320 320 > $OUT/trace_gemm_320.txt 2>&1; cut -c1-200 $OUT/trace_gemm_320.txt
timeout 120 python scripts/trace_gemm.py 262144 320 320 res > $OUT/trace_gemm_320_res.txt 2>&1; cut -c1-200 $OUT/trace_gemm_320_res.txt | tail -14
timeout 120 python scripts/trace_gemm.py 65536 640 640 > $OUT/trace_gemm_640.txt 2>&1; cut -c1-200 $OUT/trace_gemm_640.txt | tail -14

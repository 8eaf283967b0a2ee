#!/bin/bash
OUT=gpurun_out/r3q; mkdir -p $OUT
timeout 600 python -m pytest tests/test_tc_gpu.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -2
timeout 120 python scripts/trace_gemm.py 262144 320 320 > $OUT/trace_gemm_320.txt 2>&1; cut -c1-200 $OUT/trace_gemm_320.txt | tail -8

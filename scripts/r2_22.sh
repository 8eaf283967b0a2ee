#!/bin/bash
OUT=gpurun_out/r2y; mkdir -p $OUT
timeout 900 python -m pytest tests/test_tc_gpu.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -2
timeout 300 python scripts/micro.py gemm 10 2>&1 | tee $OUT/micro_gemm.txt
timeout 300 python scripts/micro.py conv 10 2>&1 | tee $OUT/micro_conv.txt

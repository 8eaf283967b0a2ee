#!/bin/bash
OUT=gpurun_out/r3n; mkdir -p $OUT
timeout 300 python scripts/micro.py ln 20 2>&1 | tee $OUT/micro_ln.txt

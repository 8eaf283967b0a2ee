#!/bin/bash
OUT=gpurun_out/r3u; mkdir -p $OUT
for v in A V0 V3 N A N; do echo "== $v"; if [ $v = N ]; then unset SIDLSG_LIB; else export SIDLSG_LIB=$PWD/sid_lsg_b200/_C/ab/lib$v.so; fi; timeout 300 python scripts/micro.py gemm 20 2>&1 | grep "linear fwd" | head -4; timeout 300 python scripts/micro.py conv 20 2>&1 | grep "conv fwd" | head -2; done 2>&1 | tee $OUT/ab_bits.txt

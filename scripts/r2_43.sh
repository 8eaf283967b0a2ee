#!/bin/bash
# which of the epilogue changes slows the conv tiles: same box, one library per A/B bit
OUT=gpurun_out/r3t; mkdir -p $OUT
for v in A V0 V8 V1 V2 V4 N A; do echo "== $v"; if [ $v = N ]; then unset SIDLSG_LIB; else export SIDLSG_LIB=$PWD/sid_lsg_b200/_C/ab/lib$v.so; fi; timeout 300 python scripts/micro.py gemm 20 2>&1 | grep "linear fwd" | head -2; timeout 300 python scripts/micro.py conv 20 2>&1 | grep "conv fwd" | head -2; done 2>&1 | tee $OUT/ab_bits.txt

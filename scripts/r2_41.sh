#!/bin/bash
# GEMM epilogue: tile cursor, hoisted fast-path predicate, one barrier per staged chunk - against the previous build (libA)
OUT=gpurun_out/r3r; mkdir -p $OUT
timeout 900 python -m pytest tests/test_tc_gpu.py tests/test_ops_gpu.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -2
timeout 120 python scripts/trace_gemm.py 262144 320 320 > $OUT/trace_gemm_320.txt 2>&1; cut -c1-200 $OUT/trace_gemm_320.txt | tail -8
for rep in 1 2; do for v in A N; do echo "== $v (rep $rep)"; if [ $v = A ]; then export SIDLSG_LIB=$PWD/sid_lsg_b200/_C/ab/libA.so; else unset SIDLSG_LIB; fi; timeout 300 python scripts/micro.py gemm 20 2>&1 | grep "linear fwd"; timeout 300 python scripts/micro.py conv 20 2>&1 | grep "conv fwd"; done; done 2>&1 | tee $OUT/ab.txt

#!/bin/bash
# same-box A/B of the GEMM epilogue changes: micro + whole-iteration graph replay
OUT=gpurun_out/r3s; mkdir -p $OUT
for rep in 1 2; do for v in A N; do echo "== $v (rep $rep)"; if [ $v = A ]; then export SIDLSG_LIB=$PWD/sid_lsg_b200/_C/ab/libA.so; else unset SIDLSG_LIB; fi; timeout 300 python scripts/micro.py gemm 20 2>&1 | grep "linear fwd" | head -4; timeout 300 python scripts/micro.py conv 20 2>&1 | grep "conv fwd" | head -3; done; done 2>&1 | tee $OUT/ab_micro.txt
for v in A N A N; do if [ $v = A ]; then export SIDLSG_LIB=$PWD/sid_lsg_b200/_C/ab/libA.so; else unset SIDLSG_LIB; fi; timeout 600 python bench.py --no-cpu-baseline --no-e2e --no-roofline --steps 4 > $OUT/bench_$v.json 2> $OUT/bench_$v.err; echo "== $v: $(python scripts/bench_brief.py $OUT/bench_$v.json | head -1)"; done 2>&1 | tee $OUT/ab_bench.txt

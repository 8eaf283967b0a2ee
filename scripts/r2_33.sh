#!/bin/bash
# GroupNorm walked in L2-sized sample groups
OUT=gpurun_out/r3j; mkdir -p $OUT
timeout 600 python -m pytest tests/test_ops_gpu.py -m gpu -q -x -p no:cacheprovider -k "norm or gn" 2>&1 | tail -2
for mb in 0 24 40 64; do echo "== SIDLSG_GN_L2MB=$mb"; SIDLSG_GN_L2MB=$mb timeout 300 python scripts/micro.py gn 20 2>&1 | tee $OUT/micro_gn_$mb.txt; done

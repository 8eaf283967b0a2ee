#!/bin/bash
# same-box A/B: A = single low-latency issuer (commit 624427d), C = forward 4 issuers + early S release, backward 2 issuers
OUT=gpurun_out/r3i; mkdir -p $OUT
for rep in 1 2 3; do for v in A C; do echo "== lib$v (rep $rep)"; SIDLSG_LIB=$PWD/sid_lsg_b200/_C/ab/lib$v.so timeout 300 python scripts/micro.py attn 20 2>&1 | grep -A1 "attn fwd" | head -2; done; done 2>&1 | tee $OUT/ab.txt

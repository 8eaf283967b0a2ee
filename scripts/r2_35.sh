#!/bin/bash
# forward attention: exponentials all on MUFU (POLY=0) against 6 of 16 pairs on the FMA pipe, same box, alternating
OUT=gpurun_out/r3l; mkdir -p $OUT
for rep in 1 2 3; do for pz in 0 6; do echo "== SIDLSG_ATTN_POLY=$pz (rep $rep)"; SIDLSG_ATTN_POLY=$pz timeout 300 python scripts/micro.py attn 20 2>&1 | grep "attn fwd" | head -3; done; done 2>&1 | tee $OUT/poly_ab.txt

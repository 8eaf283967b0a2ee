#!/bin/bash
OUT=gpurun_out/r3f; mkdir -p $OUT
for pz in 0 4 6 8; do echo "== SIDLSG_ATTN_POLY=$pz"; SIDLSG_ATTN_POLY=$pz timeout 300 python scripts/micro.py attn 10 2>&1 | grep "attn fwd" | head -2; done

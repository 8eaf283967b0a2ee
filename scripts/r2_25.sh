#!/bin/bash
OUT=gpurun_out/r3b; mkdir -p $OUT
for hn in 0 100 1000; do echo "== hint $hn"; SIDLSG_WAIT_HINT_NS=$hn timeout 120 python scripts/trace_attn_fwd.py > $OUT/trace_hint$hn.txt 2>&1; cut -c1-220 $OUT/trace_hint$hn.txt | tail -12; SIDLSG_WAIT_HINT_NS=$hn timeout 300 python scripts/micro.py attn 10 2>&1 | grep "attn fwd" | head -1; done

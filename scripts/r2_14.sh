#!/bin/bash
# forward attention hand-shake latencies: finer phase trace + suspend-hint sweep of the helper-warp mbarrier waits
OUT=gpurun_out/r2q; mkdir -p $OUT
for hn in 0 20 100; do echo "== SIDLSG_WAIT_HINT_NS=$hn"; SIDLSG_ATTN_FWD4=0 SIDLSG_WAIT_HINT_NS=$hn timeout 300 python scripts/micro.py attn 10 2>&1 | grep -A1 fwd | tee $OUT/micro_hint_$hn.txt; done
for hn in 0 100; do SIDLSG_WAIT_HINT_NS=$hn timeout 120 python scripts/trace_attn_fwd.py > $OUT/trace_fwd3_hint$hn.txt 2>&1; cut -c1-200 $OUT/trace_fwd3_hint$hn.txt | tail -12; done

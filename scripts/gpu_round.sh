#!/bin/bash
# One GPU-box visit: parity tests, bench line, ncu launch list of one step, ncu --set full of the top kernels.
# Usage (from the repo root, under gpurun):  bash scripts/gpu_round.sh [tag] [what...]   what ⊂ {tests bench list full}
TAG=${1:-run}; shift
WHAT=${*:-tests bench list full}
OUT=gpurun_out/$TAG; mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > $OUT/gpu.txt 2>&1
for w in $WHAT; do case $w in
tests) timeout 1200 python -m pytest tests -m gpu -q --maxfail=12 -p no:cacheprovider > $OUT/pytest.log 2>&1; echo "pytest exit $?" >> $OUT/pytest.log; grep -E "^(FAILED|ERROR)|passed|failed" $OUT/pytest.log | tail -20;;
smoke) timeout 600 python __graft_entry__.py smoke > $OUT/smoke.log 2>&1; tail -3 $OUT/smoke.log;;
bench) timeout 1500 python bench.py --shapes $OUT/shapes.txt > $OUT/bench.json 2> $OUT/bench.err; tail -c 3000 $OUT/bench.json;;
benchq) timeout 1500 python bench.py --no-cpu-baseline --steps 3 --shapes $OUT/shapes.txt > $OUT/bench.json 2> $OUT/bench.err; tail -c 1500 $OUT/bench.err; python scripts/bench_brief.py $OUT/bench.json;;
benchmb) timeout 900 python bench.py --no-cpu-baseline --steps 3 --batch-gpu ${MB:-32} --shapes $OUT/shapes_mb.txt > $OUT/bench_mb.json 2> $OUT/bench_mb.err; tail -c 800 $OUT/bench_mb.err; python scripts/bench_brief.py $OUT/bench_mb.json;;
benchkb) SIDLSG_BM2_MINKB=${MINKB:-8} timeout 900 python bench.py --no-cpu-baseline --steps 3 --shapes $OUT/shapes_kb.txt > $OUT/bench_kb.json 2> $OUT/bench_kb.err; tail -c 800 $OUT/bench_kb.err; python scripts/bench_brief.py $OUT/bench_kb.json;;
hint) for h in 0 100 400 2000; do echo "== SIDLSG_WAIT_HINT_NS=$h"; SIDLSG_WAIT_HINT_NS=$h timeout 300 python scripts/micro.py attn 10 2>&1 | tee -a $OUT/hint_$h.txt; done;;
gnw) for w in 4 8 12; do echo "== SIDLSG_GN_WAVES=$w"; SIDLSG_GN_WAVES=$w timeout 300 python scripts/micro.py gn 10 2>&1 | tee -a $OUT/gnw_$w.txt; done;;
ubench_mma) timeout 120 ./sid_lsg_b200/_C/ubench_mma > $OUT/ubench_mma.txt 2>&1; cat $OUT/ubench_mma.txt;;
ubench) timeout 120 ./sid_lsg_b200/_C/ubench > $OUT/ubench.txt 2>&1; cat $OUT/ubench.txt;;
list) timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/launches.csv \
        python bench.py --batch 16 --steps 1 --warmup 0 --no-e2e --no-cpu-baseline --no-roofline > $OUT/list_bench.log 2>&1
      python scripts/ncu_launches.py $OUT/launches.csv > $OUT/launches_summary.txt 2>&1; rm -f $OUT/launches.csv.keep
      tail -30 $OUT/launches_summary.txt;;
full) for k in gemm_tc_kernel attn_fwd_kernel attn_bwd_kernel; do
        timeout 900 ncu --set full --clock-control none --import-source on -k regex:$k --launch-skip ${SKIP:-40} -c 3 -f -o $OUT/$k \
          python bench.py --batch 16 --steps 1 --warmup 0 --no-cpu-baseline --no-roofline > $OUT/full_$k.log 2>&1
      done; ls -la $OUT;;
prof) for sub in ${PROF:-attn gemm}; do
        timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:gemm_tc|attn_fwd|attn_bwd|gn_' -f -o $OUT/hot_$sub \
          python scripts/prof_target.py $sub > $OUT/prof_$sub.log 2>&1; tail -2 $OUT/prof_$sub.log
        ncu -i $OUT/hot_$sub.ncu-rep --page raw --csv > $OUT/hot_${sub}_raw.csv 2>/dev/null
        python scripts/ncu_lines.py $OUT/hot_$sub.ncu-rep 70 > $OUT/hot_${sub}_lines.txt 2>&1
        sz=$(stat -c %s $OUT/hot_$sub.ncu-rep); echo "$sub report bytes $sz"
        if [ "$sz" -gt 25000000 ]; then rm -f $OUT/hot_$sub.ncu-rep; echo "(report dropped: too large to bring back)"; fi
      done; du -sh gpurun_out;;
micro) timeout 600 python scripts/micro.py ${MICRO:-all} 5 > $OUT/micro.txt 2>&1; cat $OUT/micro.txt;;
esac; done

#!/bin/bash
OUT=gpurun_out/r3w; mkdir -p $OUT
timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:gemm_tc' -f -o $OUT/hot_gemm python scripts/prof_target.py gemm > $OUT/prof_gemm.log 2>&1; tail -1 $OUT/prof_gemm.log
ncu -i $OUT/hot_gemm.ncu-rep --page raw --csv > $OUT/hot_gemm_raw.csv 2>/dev/null
python scripts/ncu_table.py $OUT/hot_gemm_raw.csv > $OUT/hot_gemm_table.txt; cut -c1-200 $OUT/hot_gemm_table.txt
python scripts/ncu_lines.py $OUT/hot_gemm.ncu-rep 40 > $OUT/hot_gemm_lines.txt 2>&1; rm -f $OUT/hot_gemm.ncu-rep

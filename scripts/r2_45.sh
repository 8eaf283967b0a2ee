#!/bin/bash
# N = 2 on one box with the final build, N = 1 on the same box right after
N=2
OUT=gpurun_out/r3v; mkdir -p $OUT
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 5 --warmup 3 --no-cpu-baseline > $OUT/bench_n2.json 2> $OUT/bench_n2.err; echo "exit $?"; grep -o '"value": [0-9.]*, "unit": "images/s", "n_gpus": [0-9]*' $OUT/bench_n2.json; grep -o '"ms_per_step": [0-9.]*' $OUT/bench_n2.json | head -1; grep -o '"sm_mhz": [0-9]*' $OUT/bench_n2.json | head -1
timeout 400 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-roofline > $OUT/bench_n1.json 2> $OUT/bench_n1.err; grep -o '"ms_per_step": [0-9.]*' $OUT/bench_n1.json | head -1; grep -o '"sm_mhz": [0-9]*' $OUT/bench_n1.json | head -1

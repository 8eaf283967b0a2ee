#!/bin/bash
N=${1:-2}
OUT=gpurun_out/r2r$N; mkdir -p $OUT
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 5 --warmup 3 --no-cpu-baseline > $OUT/bench_graph.json 2> $OUT/bench_graph.err; echo "exit $?"; grep -o '"value": [0-9.]*, "unit": "images/s", "n_gpus": [0-9]*' $OUT/bench_graph.json; grep -o '"ms_per_step": [0-9.]*' $OUT/bench_graph.json | head -2

#!/bin/bash
N=${1:-2}
OUT=gpurun_out/r2q$N; mkdir -p $OUT
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 5 --warmup 3 --no-cpu-baseline > $OUT/bench_graph.json 2> $OUT/bench_graph.err; tail -c 1500 $OUT/bench_graph.err; grep -o '"value": [0-9.]*, "unit": "images/s", "n_gpus": [0-9]*' $OUT/bench_graph.json; grep -o '"ms_per_step": [0-9.]*' $OUT/bench_graph.json | head -2
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 5 --warmup 3 --no-cpu-baseline --graph 0 --no-roofline > $OUT/bench_eager.json 2> $OUT/bench_eager.err; grep -o '"ms_per_step": [0-9.]*' $OUT/bench_eager.json | head -1
timeout 600 python -m pytest tests/test_ddp_gpu.py -m gpu -q -x -p no:cacheprovider > $OUT/pytest_ddp.log 2>&1; tail -3 $OUT/pytest_ddp.log

#!/bin/bash
# multi-GPU visit: product-level NCCL equivalence test, then the bench at N GPUs (overlapped gradient allreduce)
N=${1:-2}
OUT=gpurun_out/r2m$N; mkdir -p $OUT
nvidia-smi -L > $OUT/gpus.txt
timeout 600 python -m pytest tests/test_ddp_gpu.py -m gpu -q -x -p no:cacheprovider > $OUT/pytest_ddp.log 2>&1; tail -5 $OUT/pytest_ddp.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 --no-cpu-baseline > $OUT/bench_n$N.json 2> $OUT/bench_n$N.err; tail -c 300 $OUT/bench_n$N.err; python scripts/bench_brief.py $OUT/bench_n$N.json | head -4
timeout 600 python bench.py --gpus 1 --steps 5 --warmup 3 --no-cpu-baseline --no-roofline > $OUT/bench_n1.json 2> $OUT/bench_n1.err; python scripts/bench_brief.py $OUT/bench_n1.json | head -2

#!/bin/bash
# N GPUs: NCCL channel count against step time (NCCL CTAs take SMs away from the persistent 148-CTA kernels)
N=${1:-2}
OUT=gpurun_out/r3m$N; mkdir -p $OUT
for ch in default 2 4; do
  if [ $ch = default ]; then unset NCCL_MAX_NCHANNELS; else export NCCL_MAX_NCHANNELS=$ch; fi
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-roofline > $OUT/bench_ch$ch.json 2> $OUT/bench_ch$ch.err; echo "ch=$ch exit $?"; grep -o '"value": [0-9.]*, "unit": "images/s", "n_gpus": [0-9]*' $OUT/bench_ch$ch.json; grep -o '"ms_per_step": [0-9.]*' $OUT/bench_ch$ch.json | head -1; grep -o '"sm_mhz": [0-9]*' $OUT/bench_ch$ch.json | head -1
done
timeout 400 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-roofline > $OUT/bench_n1.json 2> $OUT/bench_n1.err; grep -o '"ms_per_step": [0-9.]*' $OUT/bench_n1.json | head -1; grep -o '"sm_mhz": [0-9]*' $OUT/bench_n1.json | head -1

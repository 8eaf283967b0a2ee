#!/bin/bash
OUT=gpurun_out/r2n; mkdir -p $OUT
timeout 600 python scripts/graph_probe.py 32 > $OUT/graph_probe.txt 2>&1; tail -12 $OUT/graph_probe.txt

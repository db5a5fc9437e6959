#!/bin/bash
# round 2, call 19 (the last GPU seconds of the round): the batched (5b) rows of the clustered generator
out=gpurun_out/r2_call19; mkdir -p $out
timeout 100 python bench_sweep.py --configs 5 --batched --generators clustered --densities 0.01,0.25 --steps 8 > $out/sweep5_clustered_batched.jsonl 2> $out/err.log
cut -c1-300 $out/sweep5_clustered_batched.jsonl

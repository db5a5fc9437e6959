#!/bin/bash
# round 2 (gpurun --gpus 4): the bench at N = 4 exactly as the driver launches it (both Count merges), wall clock of each run noted
set -u
out=gpurun_out/r2_4gpu; mkdir -p $out
nvidia-smi -L > $out/gpus.txt 2>&1
t0=$(date +%s)
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 4 --steps 20 --warmup 3 > $out/bench_n4_p2p.json 2> $out/bench_n4_p2p.err; echo "bench p2p rc=$? wall=$(( $(date +%s) - t0 ))s" >> $out/summary.txt
t0=$(date +%s)
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 4 --steps 20 --warmup 3 --reduce nccl --extras north_star,config4 > $out/bench_n4_nccl.json 2> $out/bench_n4_nccl.err; echo "bench nccl rc=$? wall=$(( $(date +%s) - t0 ))s" >> $out/summary.txt
cat $out/summary.txt; tail -c 300 $out/bench_n4_p2p.err

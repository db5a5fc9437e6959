#!/bin/bash
# round 2, call 18 (gpurun --gpus 2): the multi-GPU tests and the N = 2 bench with the config-4 sub-record after groupby_direct_kernel became
# the default GroupBy path (NCCL all-reduce of the count tensor; parity of the full tensor against the CPU port over all shards of both ranks)
set -u
out=gpurun_out/r2_call18; mkdir -p $out
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 20 --warmup 3 --extras config4 > $out/bench_n2.json 2> $out/bench_n2.err; echo "bench n2 rc=$?" >> $out/summary.txt
timeout 170 python -m pytest tests/test_gpu_multi.py tests/test_gpu_node.py -x -q -m gpu > $out/pytest_multi.log 2>&1; echo "pytest_multi rc=$?" >> $out/summary.txt
cat $out/summary.txt; tail -3 $out/pytest_multi.log; tail -c 300 $out/bench_n2.err

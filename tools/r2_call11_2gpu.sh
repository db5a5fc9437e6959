#!/bin/bash
# round 2, call 11 (gpurun --gpus 2): the multi-GPU tests (fused peer-memory Count merge, NCCL vectors, one-process node handle) and the
# bench at N = 2 with both Count merges; every sub-record carries its parity check against the CPU port over all shards of all ranks
set -u
out=gpurun_out/r2_call11; mkdir -p $out
nvidia-smi -L > $out/gpus.txt 2>&1
timeout 1200 python -m pytest tests/test_gpu_multi.py tests/test_gpu_node.py -x -q -m gpu > $out/pytest_multi.log 2>&1; echo "pytest_multi rc=$?" >> $out/summary.txt
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > $out/bench_n2_p2p.json 2> $out/bench_n2_p2p.err; echo "bench p2p rc=$?" >> $out/summary.txt
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 3 --reduce nccl --extras config4 > $out/bench_n2_nccl.json 2> $out/bench_n2_nccl.err; echo "bench nccl rc=$?" >> $out/summary.txt
timeout 600 python bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > $out/bench_ref.json 2>> $out/bench_n2_nccl.err
cat $out/summary.txt; tail -3 $out/pytest_multi.log; tail -c 600 $out/bench_n2_p2p.err

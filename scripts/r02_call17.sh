#!/bin/bash
# 2 GPUs after the chunk-loop changes: sharding tests over NCCL, then the bench line at N=2
mkdir -p gpurun_out
echo "== shard tests"; timeout 300 python -m pytest tests/test_shard_gpu.py -x -q --timeout 200 2>&1 | tail -4
echo "== bench N=2"; timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu > gpurun_out/r02_bench_2gpu_b.json 2> gpurun_out/r02_bench_2gpu_b.err; cut -c1-300 gpurun_out/r02_bench_2gpu_b.json; tail -3 gpurun_out/r02_bench_2gpu_b.err

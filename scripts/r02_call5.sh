#!/bin/bash
# 8 GPUs: sharding tests over NCCL at world 8, then the bench line at N=8 and N=4
mkdir -p gpurun_out
nvidia-smi -L | wc -l
echo "== shard tests (world 8)"; timeout 600 python -m pytest tests/test_shard_gpu.py -x -q --timeout 500 2>&1 | tail -5
for N in 8 4; do
echo "== bench N=$N"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r02_bench_${N}gpu.json 2> gpurun_out/r02_bench_${N}gpu.err; tail -c 300 gpurun_out/r02_bench_${N}gpu.json; tail -3 gpurun_out/r02_bench_${N}gpu.err
done

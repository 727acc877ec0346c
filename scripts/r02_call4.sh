#!/bin/bash
# 2 GPUs: sharding tests over NCCL, then the bench line at N=2 (sub-records with seam checks)
mkdir -p gpurun_out
nvidia-smi -L
echo "== shard tests"; timeout 900 python -m pytest tests/test_shard_gpu.py tests/test_api_gpu.py::test_filter_follows_the_device_of_its_input -x -q --timeout 600 2>&1 | tail -8
echo "== bench N=2"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02_bench_2gpu.json 2> gpurun_out/r02_bench_2gpu.err; tail -c 5000 gpurun_out/r02_bench_2gpu.json; tail -5 gpurun_out/r02_bench_2gpu.err

#!/bin/bash
# final multi-GPU record of the round: shard tests at world 8, bench lines at N = 8, 4, 2, 1
mkdir -p gpurun_out
echo "== shard tests (world 8)"; timeout 400 python -m pytest tests/test_shard_gpu.py tests/test_api_gpu.py::test_filter_follows_the_device_of_its_input -x -q --timeout 380 2>&1 | tail -3
for N in 8 4 2; do
echo "== bench N=$N"; timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2952$N bench.py --gpus $N > gpurun_out/r02f_bench_${N}gpu.json 2> gpurun_out/r02f_bench_${N}gpu.err; tail -c 200 gpurun_out/r02f_bench_${N}gpu.json; tail -2 gpurun_out/r02f_bench_${N}gpu.err | cut -c1-200
done
echo "== bench N=1"; timeout 500 python bench.py > gpurun_out/r02f_bench_1gpu.json 2> gpurun_out/r02f_bench_1gpu.err; tail -c 200 gpurun_out/r02f_bench_1gpu.json

#!/bin/bash
mkdir -p gpurun_out
echo "== full gpu tests"; timeout 1200 python -m pytest tests -m gpu -x -q --timeout 600 --durations=5 > gpurun_out/r02_pytest7.log 2>&1; tail -10 gpurun_out/r02_pytest7.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== configs c2 c3 c5"; timeout 600 python scripts/bench_configs.py --only c2,c3,c5 > gpurun_out/r02_bench_configs_c2c3c5.jsonl 2> gpurun_out/r02_bench_configs_c2c3c5.err; cut -c1-700 gpurun_out/r02_bench_configs_c2c3c5.jsonl; tail -3 gpurun_out/r02_bench_configs_c2c3c5.err
echo "== bench N=1"; timeout 900 python bench.py > gpurun_out/r02_bench_1gpu_b.json 2> gpurun_out/r02_bench_1gpu_b.err; cut -c1-900 gpurun_out/r02_bench_1gpu_b.json; tail -3 gpurun_out/r02_bench_1gpu_b.err

#!/bin/bash
mkdir -p gpurun_out
echo "== full gpu tests"; timeout 1200 python -m pytest tests -m gpu -x -q --timeout 600 --durations=6 > gpurun_out/r02_pytest3.log 2>&1; tail -14 gpurun_out/r02_pytest3.log
echo "== ops"; timeout 600 python scripts/bench_ops.py > gpurun_out/r02_bench_ops.jsonl 2> gpurun_out/r02_bench_ops.err; cat gpurun_out/r02_bench_ops.jsonl | cut -c1-260; tail -3 gpurun_out/r02_bench_ops.err
echo "== configs c1 c2"; timeout 600 python scripts/bench_configs.py --only c1,c2 2>&1 | cut -c1-600
echo "== reference arm"; timeout 600 python bench.py --impl reference --steps 2 --warmup 1 | cut -c1-900
echo "== bench N=1"; timeout 900 python bench.py > gpurun_out/r02_bench_1gpu.json 2> gpurun_out/r02_bench_1gpu.err; cut -c1-700 gpurun_out/r02_bench_1gpu.json; tail -3 gpurun_out/r02_bench_1gpu.err

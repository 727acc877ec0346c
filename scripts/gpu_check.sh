#!/bin/bash
# Runs on the GPU box (via gpurun): smoke, GPU tests, a short bench, ncu launch list.  Output -> gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
echo "== smoke" ; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5
echo "== pytest -m gpu" ; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -25
echo "== bench" ; timeout 900 python bench.py --steps 10 --warmup 3 --cpu-reps 1 > gpurun_out/bench.json 2> gpurun_out/bench.err
tail -c 3000 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
echo "== ncu launches" ; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu --e2e-samples 40000000 > gpurun_out/ncu_bench.log 2>&1
tail -5 gpurun_out/launches.csv

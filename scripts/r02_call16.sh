#!/bin/bash
mkdir -p gpurun_out
echo "== ncu launch list of the bench command"
timeout 240 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu --e2e-samples 40000000 > gpurun_out/r02_ncu_bench.log 2>&1
echo "rc=$?"; wc -l gpurun_out/r02_launches_bench.csv; tail -3 gpurun_out/r02_ncu_bench.log | cut -c1-300

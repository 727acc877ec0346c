#!/bin/bash
mkdir -p gpurun_out
echo "== configs c2"; timeout 600 python scripts/bench_configs.py --only c2 > gpurun_out/r02_bench_configs_c2.jsonl 2> gpurun_out/r02_bench_configs_c2.err; cut -c1-900 gpurun_out/r02_bench_configs_c2.jsonl; tail -3 gpurun_out/r02_bench_configs_c2.err

#!/bin/bash
mkdir -p gpurun_out
echo "== sync/image tests"; timeout 600 python -m pytest tests/test_sync_gpu.py -m "gpu and not slow" -x -q --timeout 300 > gpurun_out/r02_pytest8.log 2>&1; tail -4 gpurun_out/r02_pytest8.log
echo "== configs c2 c3"; timeout 600 python scripts/bench_configs.py --only c2,c3 > gpurun_out/r02_bench_configs_c2c3.jsonl 2> gpurun_out/r02_bench_configs_c2c3.err; cut -c1-1000 gpurun_out/r02_bench_configs_c2c3.jsonl; tail -3 gpurun_out/r02_bench_configs_c2c3.err

#!/bin/bash
mkdir -p gpurun_out
echo "== quick tests"; timeout 600 python -m pytest tests/test_api_gpu.py tests/test_chain_gpu.py -x -q --timeout 300 2>&1 | tail -3
echo "== chain A/B"; timeout 600 python scripts/chain_ab.py 1843200000 d34 2>&1 | cut -c1-200
echo "== bench"; timeout 600 python bench.py --steps 20 --warmup 3 --no-extra --cpu-reps 1 2> gpurun_out/r02_bench3.err | tee gpurun_out/r02_bench3.json | cut -c1-1800

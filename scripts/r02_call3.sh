#!/bin/bash
mkdir -p gpurun_out
echo "== chain tests"; timeout 600 python -m pytest tests/test_chain_gpu.py tests/test_api_gpu.py -x -q --timeout 300 2>&1 | tail -5
echo "== chain A/B"; timeout 900 python scripts/chain_ab.py > gpurun_out/r02_chain_ab2.jsonl 2> gpurun_out/r02_chain_ab2.err; cat gpurun_out/r02_chain_ab2.jsonl; tail -5 gpurun_out/r02_chain_ab2.err

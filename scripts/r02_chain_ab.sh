#!/bin/bash
mkdir -p gpurun_out
echo "== readbw2"; timeout 300 ./scratch/readbw2 > gpurun_out/r02_readbw2.txt 2>&1; cat gpurun_out/r02_readbw2.txt
echo "== chain tests with the new kernel"; timeout 900 python -m pytest tests/test_chain_gpu.py tests/test_api_gpu.py -x -q 2>&1 | tail -8
echo "== chain A/B"; timeout 900 python scripts/chain_ab.py > gpurun_out/r02_chain_ab.jsonl 2> gpurun_out/r02_chain_ab.err; cat gpurun_out/r02_chain_ab.jsonl; tail -5 gpurun_out/r02_chain_ab.err

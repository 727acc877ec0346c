#!/bin/bash
mkdir -p gpurun_out
echo "== readbw2"; timeout 200 ./scratch/readbw2 > gpurun_out/r02_readbw2.txt 2>&1; cat gpurun_out/r02_readbw2.txt
echo "== chain A/B"; timeout 600 python scripts/chain_ab.py > gpurun_out/r02_chain_ab.jsonl 2> gpurun_out/r02_chain_ab.err; cat gpurun_out/r02_chain_ab.jsonl; tail -5 gpurun_out/r02_chain_ab.err
echo "== iir sweep"; timeout 300 python scripts/iir_sweep.py > gpurun_out/r02_iir_sweep.jsonl 2>&1; cat gpurun_out/r02_iir_sweep.jsonl
echo "== tests"; timeout 1500 python -m pytest tests/test_chain_gpu.py tests/test_api_gpu.py tests/test_c1_gpu.py tests/test_sync_gpu.py -x -q --timeout 600 --durations=15 > gpurun_out/r02_pytest2.log 2>&1; tail -40 gpurun_out/r02_pytest2.log

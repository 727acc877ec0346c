#!/bin/bash
mkdir -p gpurun_out
echo "== chain + api + shard tests"; timeout 600 python -m pytest tests/test_chain_gpu.py tests/test_api_gpu.py tests/test_shard_gpu.py -m gpu -x -q --timeout 300 > gpurun_out/r02_pytest5.log 2>&1; tail -5 gpurun_out/r02_pytest5.log
echo "== audio breakdown"; timeout 300 python scripts/audio_breakdown.py > gpurun_out/r02_audio_breakdown2.jsonl 2> gpurun_out/r02_audio_breakdown2.err; cat gpurun_out/r02_audio_breakdown2.jsonl; tail -3 gpurun_out/r02_audio_breakdown2.err

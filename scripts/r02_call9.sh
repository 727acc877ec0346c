#!/bin/bash
mkdir -p gpurun_out
echo "== chain + api tests"; timeout 600 python -m pytest tests/test_chain_gpu.py tests/test_api_gpu.py tests/test_shard_gpu.py -m gpu -x -q --timeout 300 > gpurun_out/r02_pytest4.log 2>&1; tail -5 gpurun_out/r02_pytest4.log
echo "== audio breakdown (halo carried in the kernel)"; timeout 300 python scripts/audio_breakdown.py --profile > gpurun_out/r02_audio_breakdown.jsonl 2> gpurun_out/r02_audio_breakdown_cprofile.txt; cat gpurun_out/r02_audio_breakdown.jsonl; head -45 gpurun_out/r02_audio_breakdown_cprofile.txt | cut -c1-150
echo "== audio breakdown (halo carried by a copy node: previous behaviour)"; DDM_CHAIN_HALO_MEMCPY=1 timeout 300 python scripts/audio_breakdown.py >> gpurun_out/r02_audio_breakdown.jsonl 2> gpurun_out/r02_audio_breakdown.err; tail -1 gpurun_out/r02_audio_breakdown.jsonl; tail -3 gpurun_out/r02_audio_breakdown.err

#!/bin/bash
# ncu --set full capture of the fused chain kernel at the bench workload (1 GPU).
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:chain_fused -s 3 -c 2 \
    -o gpurun_out/prof_chain -f python bench.py --steps 2 --warmup 3 --no-cpu --e2e-samples 20000000 --e2e-steps 1 > gpurun_out/prof_chain.log 2>&1
tail -3 gpurun_out/prof_chain.log
ls -la gpurun_out

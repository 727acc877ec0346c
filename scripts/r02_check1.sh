#!/bin/bash
# round 2, first GPU pass: smoke, GPU tests, full bench line, ncu full capture (with warp states) of the fused chain.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt; free -g >> gpurun_out/gpu.txt
echo "== smoke" ; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5
echo "== pytest -m gpu" ; timeout 2400 python -m pytest tests -m gpu -x -q --durations=12 2>&1 | tail -40 | tee gpurun_out/r02_pytest1.log
echo "== bench" ; timeout 1200 python bench.py > gpurun_out/r02_bench1.json 2> gpurun_out/r02_bench1.err
tail -c 6000 gpurun_out/r02_bench1.json; tail -5 gpurun_out/r02_bench1.err
echo "== ncu full chain"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:chain_fused -s 3 -c 1 \
    -o gpurun_out/r02_prof_chain0 -f python bench.py --steps 2 --warmup 3 --no-cpu --no-u8 --no-extra --e2e-samples 20000000 --e2e-steps 1 > gpurun_out/r02_prof_chain0.log 2>&1
tail -2 gpurun_out/r02_prof_chain0.log
ls -la gpurun_out | tail -8

#!/bin/bash
mkdir -p gpurun_out
echo "== quick tests"; timeout 600 python -m pytest tests/test_api_gpu.py tests/test_chain_gpu.py -x -q --timeout 300 2>&1 | tail -3
echo "== iir sweep"; timeout 200 python scripts/iir_sweep.py > gpurun_out/r02_iir_sweep2.jsonl 2>&1; cat gpurun_out/r02_iir_sweep2.jsonl
echo "== configs"; timeout 900 python scripts/bench_configs.py > gpurun_out/r02_bench_configs.jsonl 2> gpurun_out/r02_bench_configs.err; cat gpurun_out/r02_bench_configs.jsonl; tail -3 gpurun_out/r02_bench_configs.err
echo "== python profile"; timeout 300 python scripts/prof_python.py > gpurun_out/r02_prof_python.txt 2>&1; head -120 gpurun_out/r02_prof_python.txt
echo "== ncu chain"; timeout 600 ncu --set full --clock-control none --import-source on -k regex:chain_stream -s 3 -c 1 -o gpurun_out/r02_prof_chain1 -f python bench.py --steps 2 --warmup 3 --no-cpu --no-u8 --no-extra --e2e-samples 20000000 --e2e-steps 1 > gpurun_out/r02_prof_chain1.log 2>&1; tail -2 gpurun_out/r02_prof_chain1.log
echo "== ncu c4 cascade"; timeout 600 ncu --set full --clock-control none --import-source on -k regex:fir_fft -s 1 -c 1 -o gpurun_out/r02_prof_c4 -f python scripts/bench_configs.py --only c4 > gpurun_out/r02_prof_c4.log 2>&1; tail -2 gpurun_out/r02_prof_c4.log

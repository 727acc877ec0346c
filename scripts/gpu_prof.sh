#!/bin/bash
# ncu captures behind the numbers in DESIGN.md / bench.py (1 GPU).  Output -> gpurun_out/.
mkdir -p gpurun_out
# (1) every launch of the default bench command with its device time
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu --e2e-samples 40000000 --e2e-steps 1 > gpurun_out/launches.log 2>&1
# (2) full capture of the fused chain kernel at the bench workload
timeout 900 ncu --set full --clock-control none --import-source on -k regex:chain_fused -s 3 -c 1 \
    -o gpurun_out/prof_chain -f python bench.py --steps 2 --warmup 3 --no-cpu --e2e-samples 20000000 --e2e-steps 1 > gpurun_out/prof_chain.log 2>&1
# (3) full capture of one launch of each other hot-path kernel
timeout 1200 ncu --set full --clock-control none --import-source on \
    -k regex:'fir_kernel|iir_kernel|fft_pass|ncc_tile_kernel|mix_kernel|fm_kernel|select_hist2|compact_write|bank4|hilbert_mask|abs_out' \
    -s 24 -c 30 -o gpurun_out/prof_ops -f python scripts/prof_ops.py > gpurun_out/prof_ops.log 2>&1
tail -2 gpurun_out/prof_chain.log gpurun_out/prof_ops.log
ls -la gpurun_out

#!/bin/bash
# ncu --set full of one launch of each non-FFT hot-path kernel (scripts/prof_ops.py, second round)
mkdir -p gpurun_out
timeout 1500 ncu --set full --clock-control none --import-source on \
    -k regex:'fir_kernel|iir_kernel|ncc_tile_kernel|mix_kernel|fm_kernel|select_hist2|compact_write|bank4|row_median|chain_fused' \
    -s 16 -c 16 -o gpurun_out/prof_ops2 -f python scripts/prof_ops.py > gpurun_out/prof_ops2.log 2>&1
tail -n 2 gpurun_out/prof_ops2.log
# the C4 kernels of the second half of round 1 (scripts/prof_c4.py)
timeout 600 ncu --set full --clock-control none --import-source on \
    -k regex:'fir_fft_kernel|iir_warp_kernel|iir_seq_warp' -s 4 -c 4 -o gpurun_out/prof_c4 -f python scripts/prof_c4.py > gpurun_out/prof_c4.log 2>&1
tail -n 2 gpurun_out/prof_c4.log

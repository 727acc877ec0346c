#!/usr/bin/env python
"""One launch of the C4 kernels (overlap-save FFT FIR, warp-staged IIR, sequential warp replay) --
the target of the ncu --set full capture summarised in profiles/r01_c4_kernels_ncu_full.csv
(scripts/gpu_prof_ops.sh)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from directdemod_b200 import filters, constants
torch.cuda.set_device(0)
n = 600_000_000          # above the size where the IIR keeps one complex recursion per lane
x = torch.empty(n, dtype=torch.complex64, device="cuda")
torch.view_as_real(x).normal_(0, 40)
xr = torch.empty(400_000, dtype=torch.float32, device="cuda").normal_()
fir = filters.remez(2400000, [[0, 100000], [120000, 1199999]], [1, 0], ntaps=1023)
iir = filters.butter(2400000, 100000, n=8)
bp = filters.butter(60235, 400, 4400, n=6, typeFlt=constants.FLT_BP)
for rep in range(2):
    fir._apply_dev(x)
    iir._apply_dev(x)
    iir._apply_dev(x[:100_000_000])      # chunk-sized: re/im split over lane pairs
    bp._apply_dev(xr)
    torch.cuda.synchronize()
print("done")

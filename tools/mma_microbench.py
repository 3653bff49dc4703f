#!/usr/bin/env python
"""tcgen05.mma issue-rate micro-benchmark (M=128, K=16, A in TMEM): cycles per MMA for dependent chains vs
round-robin over disjoint accumulators, by N.  Run on the GPU box: python tools/mma_microbench.py"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from nphm_b200 import _native
lib = _native.lib()
lib.nphm_debug_tc_mma_bench.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_longlong)]
torch.cuda.init(); torch.zeros(1, device='cuda')
print('%5s %6s %10s %12s %10s' % ('N', 'alt', 'iters', 'cyc/MMA', 'floor'))
for n in (16, 32, 48, 64, 96, 112, 208):
    for alt in (1, 2, 4):
        if alt * n > 416:
            continue
        for iters in (64, 512):
            c = ctypes.c_longlong(0)
            _native.check(lib.nphm_debug_tc_mma_bench(n, iters, alt, ctypes.byref(c)))
            if iters == 64:
                c64 = c.value
            else:
                print('%5d %6d %10d %12.1f %10.1f' % (n, alt, iters, (c.value - c64) / (512 - 64), 128 * n / 256))

#!/usr/bin/env python
"""Timeline of the dense ensemble kernel (v8) from its built-in trace points.

    tools/build_variant.sh trace -DNPHM_TC_TRACE
    NPHM_B200_LIB=$PWD/nphm_b200/libnphm_b200_trace.so python tools/tc_trace.py > gpurun_out/tc_trace.txt

CTA 0 stamps clock64() at the phase boundaries of its second tile; printed per member, in SM cycles relative to the
start of member 4's iteration on compute warp 0."""
import ctypes
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
from conftest import MAXI, MINI, make_ensemble, sample_latent       # noqa: E402
from nphm_b200 import _native                                       # noqa: E402

EP = ['iter start', 'E0b done', 'E3b(prev)+blend done', 'M1 done seen', 'E1 done', 'M2 done seen', 'E2 done',
      'waiting for M3', 'M3 done seen', 'E3a done', 'E0a(next) done', 'loop top (before the record wait)']
IS = {16: 'M1 may start', 17: 'M1 S k-steps issued', 18: 'M1 all issued', 19: 'a1[0] seen', 20: 'a1[1] seen',
      22: 'M2 all issued', 23: 'a2[0] seen', 24: 'a2[1] seen', 25: 'a2[2] seen', 26: 'a2[3] seen', 28: 'M3 all issued'}


def main():
    dev = torch.device('cuda', 0)
    dec = make_ensemble(0, device=dev).eval()
    lat = sample_latent(1).to(dev)
    res = 128
    for _ in range(2):
        dec.engine().query_grid(lat, MINI, MAXI, res, 0, res ** 3, quirk_period=25000)
    torch.cuda.synchronize()
    buf = np.zeros((64, 48), np.int64)
    lib = _native.lib()
    rc = lib.nphm_debug_tc_trace(ctypes.c_void_p(buf.ctypes.data), ctypes.c_int(buf.size))
    assert rc == 0
    t0 = buf[4, 0]
    for m in range(4, 9):
        print('--- member %d (cycles since member 4 started; member length %d)' % (m, buf[m + 1, 0] - buf[m, 0]))
        ev = []
        for i, name in enumerate(EP):
            ev.append((buf[m, i] - t0, 'warp0  ' + name))
            ev.append((buf[m, 32 + i] - t0, 'warp13 ' + name))
        for i, name in IS.items():
            ev.append((buf[m, i] - t0, 'ISSUER ' + name))
        for t, name in sorted(ev):
            print('%8d  %s' % (t, name))


if __name__ == '__main__':
    main()

#!/usr/bin/env python
"""SS-form tcgen05.mma probe: variant bits: 0-1 A layout (0 chunk-major, 2 slab-like), 4 swap descriptor order, 8 M=128."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from nphm_b200 import _native
lib = _native.lib()
lib.nphm_debug_tc_mma_m64.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                      ctypes.c_void_p, ctypes.c_void_p]
rng = np.random.RandomState(0)
n, ks = 128, 2
K = 16 * ks
for variant in (0, 2, 4, 6, 8, 10, 12, 14):
    m = 128 if variant & 8 else 64
    a = (rng.randn(m, K) * 2).astype(np.float32)
    b = (rng.randn(n, K) * 0.5).astype(np.float32)
    lanes = np.arange(128) if m == 128 else np.array([(r % 16) + 32 * (r // 16) for r in range(64)])
    dump = torch.zeros(128, n, device='cuda')
    ad, bd = torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()          # keep alive: .data_ptr() of a temporary dangles
    _native.check(lib.nphm_debug_tc_mma_m64(ad.data_ptr(), bd.data_ptr(), n, ks, variant, dump.data_ptr(), None))
    d = dump.cpu().numpy()[lanes].astype(np.float64)
    cands = {'A@B.T': a @ b.T, 'B[:m]@B.T': b[:m] @ b.T, 'A@A.T': np.pad(a @ a.T, ((0, 0), (0, n - m)))[:, :n] if m <= n else None,
             'B[:m]@A.T': np.pad(b[:m] @ a.T, ((0, 0), (0, n - m)))[:, :n]}
    print('variant %2d (M=%d layout %d swap %d):' % (variant, m, variant & 3, (variant >> 2) & 1),
          {k: round(float(np.abs(d[:, :min(n, m) if 'A.T' in k else n] - v[:, :min(n, m) if 'A.T' in k else n]).max()), 5) for k, v in cands.items() if v is not None})

#!/usr/bin/env python
"""Timeline of one tc_linear CTA (build with -DNPHM_TCL_TRACE: tools/build_variant.sh): runs one layer-chain pass of the
deformation network and prints, per k-step, when the producer / issuer / row warp 0 passed their waits (cycles from CTA start).

    bash tools/build_variant.sh /tmp/tcl -DNPHM_TCL_TRACE && NPHM_B200_LIB=/tmp/tcl/libnphm_b200.so python tools/tcl_trace.py"""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np
import torch


def main():
    from conftest import make_deformation
    from nphm_b200 import _native
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 5000
    dev = torch.device('cuda', 0)
    dfn = make_deformation(device=dev)
    eng = dfn.defDeepSDF.engine()
    xyz = torch.randn(1, n, 3, device=dev) * 0.2
    cond = torch.randn(1, dfn.defDeepSDF.lat_dim, device=dev) * 0.1
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(3):
        eng.query_layers(xyz, cond)
    ev0.record()
    for _ in range(20):
        eng.query_layers(xyz, cond)
    ev1.record(); torch.cuda.synchronize()
    print('query_layers (%d points): %.1f us per call' % (n, ev0.elapsed_time(ev1) * 1000 / 20))
    g = torch.randn(1, n, 3, device=dev)
    ev0.record()
    for _ in range(20):
        eng.inverse_jacobian(xyz, cond)
    ev1.record(); torch.cuda.synchronize()
    print('inverse_jacobian: %.1f us per call' % (ev0.elapsed_time(ev1) * 1000 / 20))
    ev0.record()
    for _ in range(20):
        eng.backward_inputs(xyz, cond, g, want_xyz=True)
    ev1.record(); torch.cuda.synchronize()
    print('backward_inputs: %.1f us per call' % (ev0.elapsed_time(ev1) * 1000 / 20))
    eng.query_layers(xyz, cond)
    lib = _native.lib()
    if not hasattr(lib, 'nphm_debug_tcl_trace'):
        return
    buf = (ctypes.c_longlong * (16 * 64))()
    lib.nphm_debug_tcl_trace(buf, 16 * 64)
    t = np.array(buf[:], dtype=np.int64).reshape(16, 64)
    t0 = t[10, 0]
    names = ['prod:empty', 'iss:top', 'iss:a_full', 'iss:b_full', 'row:top', 'row:cp_done', 'row:split', 'row:empty', 'row:arrived']
    print('CTA start -> after alloc/sync: %d cycles' % (t[10, 1] - t0))
    print('%4s ' % 'j' + ' '.join('%11s' % s for s in names))
    for j in range(40):
        if t[3, j] == 0:
            break
        print('%4d ' % j + ' '.join('%11d' % (t[f, j] - t0) for f in range(9)))
    print('epilogue units of thread 0 (cycles from CTA start): top, accumulator loaded, computed, packed stored, derivative stored')
    for u in range(0, 16, 2):
        if t[11, u]:
            print('  unit %2d ' % u + ' '.join('%8d' % (t[f, u] - t0) for f in (11, 12, 13, 14, 15)))
    print('epilogue: wait d_ready from %d to %d, done %d' % (t[9, 0] - t0, t[9, 1] - t0, t[9, 2] - t0))


if __name__ == '__main__':
    main()

"""GPU parity of the evaluation metrics: nphm_nearest_neighbors / the mirror of NPHM.evaluation.metrics against
scipy's cKDTree - the library call the reference itself makes (src/NPHM/evaluation/metrics.py:184-185) - on fp32 clouds."""
import time

import numpy as np
import pytest
import torch
from scipy.spatial import cKDTree

pytestmark = pytest.mark.gpu


def _clouds(n_a, n_b, seed):
    rng = np.random.RandomState(seed)
    a = (rng.randn(n_a, 3) * 0.1).astype(np.float32)
    b = (a[rng.randint(0, n_a, n_b)] + rng.randn(n_b, 3).astype(np.float32) * 0.004).astype(np.float32)
    return a, b


def test_nearest_neighbors_match_ckdtree(cuda_device):
    from nphm_b200 import _native
    for n_a, n_b in ((1, 7), (1000, 37), (4099, 5001)):
        a, b = _clouds(n_a, n_b, n_a)
        dist, idx = _native.nearest_neighbors(torch.from_numpy(b).to(cuda_device), torch.from_numpy(a).to(cuda_device))
        d_ref, i_ref = cKDTree(a.astype(np.float64)).query(b.astype(np.float64))
        assert np.abs(dist.cpu().numpy() - d_ref).max() < 1e-6
        same = idx.cpu().numpy() == i_ref
        assert same.mean() > 0.999                      # ties at fp32 round-off may pick the other point (same distance)


def test_eval_pointcloud_matches_reference_formulas(cuda_device):
    """The mirror's dictionary against the reference's formulas evaluated with cKDTree (metrics.py:46-145, metric_space=False)."""
    from nphm_b200.evaluation.metrics import eval_pointcloud
    pred, gt = _clouds(20000, 18000, 3)
    rng = np.random.RandomState(9)
    n_pred, n_gt = rng.randn(*pred.shape), rng.randn(*gt.shape)
    out, pcs = eval_pointcloud(pred.copy(), gt.copy(), n_pred, n_gt, return_error_pcs=True, metric_space=False)

    def p2p(src, tgt, ns, nt):
        d, i = cKDTree(tgt.astype(np.float64)).query(src.astype(np.float64))
        ns = ns / np.linalg.norm(ns, axis=-1, keepdims=True); nt = nt / np.linalg.norm(nt, axis=-1, keepdims=True)
        return d, np.abs((nt[i] * ns).sum(-1))
    comp, comp_n = p2p(gt, pred, n_gt, n_pred)
    acc, acc_n = p2p(pred, gt, n_pred, n_gt)
    th = [0.005, 0.01, 0.015, 0.02]
    rec = [(comp <= t).mean() for t in th]; prec = [(acc <= t).mean() for t in th]
    want = {'completeness': comp.mean(), 'accuracy': acc.mean(), 'completeness2': (comp ** 2).mean(), 'accuracy2': (acc ** 2).mean(),
            'chamfer_l1': 0.5 * (comp.mean() + acc.mean()), 'chamfer_l2': 0.5 * ((comp ** 2).mean() + (acc ** 2).mean()),
            'normals consistency': 0.5 * comp_n.mean() + 0.5 * acc_n.mean(),
            'f_score_05': 2 * prec[0] * rec[0] / (prec[0] + rec[0]), 'f_score_20': 2 * prec[3] * rec[3] / (prec[3] + rec[3])}
    for k, v in want.items():
        assert abs(out[k] - v) <= 1e-6 * max(1.0, abs(v)) + 2e-4 * (k.startswith('f_score') or k.startswith('normals')), (k, out[k], v)
    assert np.abs(pcs['completeness'] - comp).max() < 1e-6 and np.abs(pcs['accuracy'] - acc).max() < 1e-6


def test_full_size_clouds_250k(cuda_device):
    """eval.py:111 samples 250 000 points per cloud: time the GPU path next to cKDTree (printed, not asserted) and compare."""
    from nphm_b200 import _native
    a, b = _clouds(250000, 250000, 11)
    ta, tb = torch.from_numpy(a).to(cuda_device), torch.from_numpy(b).to(cuda_device)
    _native.nearest_neighbors(tb[:1000], ta)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    dist, _ = _native.nearest_neighbors(tb, ta)
    torch.cuda.synchronize(); t_gpu = time.perf_counter() - t0
    t0 = time.perf_counter()
    d_ref, _ = cKDTree(a.astype(np.float64)).query(b.astype(np.float64))
    t_cpu = time.perf_counter() - t0
    print('250k x 250k nearest neighbours: GPU %.1f ms, cKDTree %.0f ms' % (1e3 * t_gpu, 1e3 * t_cpu))
    assert np.abs(dist.cpu().numpy() - d_ref).max() < 1e-6

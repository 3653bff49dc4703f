"""Drop-in mirror of the point-cloud metrics of ``NPHM.evaluation.metrics``.

  * ``distance_p2p``              src/NPHM/evaluation/metrics.py:171-194  (cKDTree query -> ``nphm_nearest_neighbors``)
  * ``get_threshold_percentage``  src/NPHM/evaluation/metrics.py:197-208
  * ``eval_pointcloud``           src/NPHM/evaluation/metrics.py:46-145

Same return dictionaries and key names.  ``metric_space=True`` of the reference multiplies both clouds by a per-scan scale
that it reads from the dataset (``DataManager().get_transform_from_metric``); the dataset is out of scope here, so the scale is
passed in (``scale_nphm_2_metric``) - everything after that line is identical.  Inputs may be numpy arrays (uploaded) or CUDA
tensors; like the reference, ``eval_pointcloud`` with ``metric_space=True`` scales the caller's numpy arrays in place.
"""
from __future__ import annotations

import numpy as np
import torch

from .. import _native


def _device():
    if not torch.cuda.is_available():
        raise _native.NativeError('nphm_b200.evaluation.metrics needs a CUDA device (no CPU fallback)')
    return torch.device('cuda', torch.cuda.current_device())


def _to_dev(a):
    if isinstance(a, torch.Tensor):
        return a.to(device=_device() if not a.is_cuda else a.device, dtype=torch.float32)
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(_device())


def distance_p2p(pointcloud_pred, pointcloud_gt, normals_pred, normals_gt):
    """Distance of every point of ``pointcloud_pred`` to its nearest neighbour in ``pointcloud_gt`` (numpy float64) and, when
    normals are given, the |dot product| of the unit normals at those pairs."""
    pred, gt = _to_dev(pointcloud_pred), _to_dev(pointcloud_gt)
    dist, idx = _native.nearest_neighbors(pred, gt)
    dist_np = dist.cpu().numpy()
    if normals_pred is None:
        return dist_np, None
    npred = np.asarray(normals_pred) / np.linalg.norm(np.asarray(normals_pred), axis=-1, keepdims=True)
    ngt = np.asarray(normals_gt) / np.linalg.norm(np.asarray(normals_gt), axis=-1, keepdims=True)
    dots = np.abs((ngt[idx.cpu().numpy()] * npred).sum(axis=-1))         # wrong-way normals are tolerated, as upstream
    return dist_np, dots


def get_threshold_percentage(dist, thresholds):
    """Fraction of the distances below each threshold."""
    return [(dist <= t).mean() for t in thresholds]


def eval_pointcloud(pointcloud_pred, pointcloud_gt, normals_pred=None, normals_gt=None, return_error_pcs=False,
                    metric_space=True, subject=None, expression=None, scale_nphm_2_metric=None):
    """Completeness / accuracy / Chamfer-L1 / Chamfer-L2 / F-scores / normal consistency, keys as in the reference."""
    thresholds = [1, 5, 10, 20] if metric_space else [0.005, 0.01, 0.015, 0.02]
    pointcloud_pred = np.asarray(pointcloud_pred)
    pointcloud_gt = np.asarray(pointcloud_gt)
    if metric_space:
        if scale_nphm_2_metric is None:
            raise ValueError('metric_space=True needs scale_nphm_2_metric (the reference reads 1/s from its DataManager)')
        pointcloud_pred *= scale_nphm_2_metric
        pointcloud_gt *= scale_nphm_2_metric
    completeness_pc, completeness_pc_normals = distance_p2p(pointcloud_gt, pointcloud_pred, normals_gt, normals_pred)
    recall = get_threshold_percentage(completeness_pc, thresholds)
    completeness, completeness2 = completeness_pc.mean(), (completeness_pc ** 2).mean()
    accuracy_pc, accuracy_pc_normals = distance_p2p(pointcloud_pred, pointcloud_gt, normals_pred, normals_gt)
    precision = get_threshold_percentage(accuracy_pc, thresholds)
    accuracy, accuracy2 = accuracy_pc.mean(), (accuracy_pc ** 2).mean()
    F = [2 * precision[i] * recall[i] / (precision[i] + recall[i]) for i in range(len(precision))]
    if normals_pred is not None:
        accuracy_normals = accuracy_pc_normals.mean()
        completeness_normals = completeness_pc_normals.mean()
        normals_correctness = 0.5 * completeness_normals + 0.5 * accuracy_normals
    else:
        accuracy_normals = completeness_normals = normals_correctness = np.nan
    out_dict = {
        'completeness': completeness, 'accuracy': accuracy,
        'normals completeness': completeness_normals, 'normals accuracy': accuracy_normals,
        'normals consistency': normals_correctness,
        'completeness2': completeness2, 'accuracy2': accuracy2,
        'chamfer_l2': 0.5 * completeness2 + 0.5 * accuracy2, 'chamfer_l1': 0.5 * (completeness + accuracy),
        'f_score_05': F[0], 'f_score_10': F[1], 'f_score_15': F[2], 'f_score_20': F[3],
    }
    if return_error_pcs:
        return out_dict, {'completeness': completeness_pc, 'accuracy': accuracy_pc,
                          'completeness_normals': completeness_pc_normals, 'accuracy_normals': accuracy_pc_normals}
    return out_dict

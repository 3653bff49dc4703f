"""Drop-in mirror of the reference module ``NPHM.models.reconstruction``.

  * ``get_logits``           src/NPHM/models/reconstruction.py:6-25
  * ``get_logits_backward``  src/NPHM/models/reconstruction.py:28-56
  * ``deform_mesh``          src/NPHM/models/reconstruction.py:59-88

The reference evaluates the decoder chunk by chunk (``nbatch_points``) to bound its activation memory and
copies every chunk to the host.  The fused kernels keep activations on chip, so one launch covers all
points; the only chunk-size dependence of the reference's RESULT - in eval mode the last point of every
chunk gets s_k = 1 (EnsembledDeepSDF.py:260-261) - is reproduced through ``quirk_period``.
"""
from __future__ import annotations

import numpy as np
import torch

from .EnsembledDeepSDF import FastEnsembleDeepSDFMirrored
from ..utils.mesh import make_mesh
from ..utils.reconstruction import _to_host


def _as_row(encoding: torch.Tensor) -> torch.Tensor:
    """1-D ``(D,)`` or ``(1,1,D)`` latent -> ``(1, D)``."""
    return encoding.reshape(1, -1)


def get_logits(decoder, encoding, grid_points, nbatch_points=100000, return_anchors=False):
    """SDF of ``grid_points`` (1 x N x 3) under latent ``encoding`` -> numpy (N,) float32 [, anchors]."""
    with torch.no_grad():
        if isinstance(decoder, FastEnsembleDeepSDFMirrored) and grid_points.is_cuda \
                and grid_points.dtype == torch.float32:
            pts = grid_points.reshape(1, -1, 3)
            period = 0 if decoder.training else int(nbatch_points)
            sdf, anchors = decoder.engine().query(pts, _as_row(encoding).to(pts.device), eval_quirk=not decoder.training,
                                                  quirk_period=period if period else None)
            logits = _to_host(sdf.reshape(-1))
        else:
            outs = []
            anchors = None
            enc = encoding.reshape(1, 1, -1)
            for points in torch.split(grid_points, nbatch_points, dim=1):
                out, anchors = decoder(points, enc.expand(1, points.shape[1], enc.shape[-1]), None)
                outs.append(out.reshape(-1).detach().cpu())
            logits = torch.cat(outs, dim=0).numpy()
    if return_anchors:
        return logits, anchors
    return logits


def get_logits_backward(decoder_shape, decoder_expr, encoding_shape, encoding_expr, grid_points,
                        nbatch_points=100000, return_anchors=False):
    """Backward-warp query: points are first offset by ``decoder_expr`` (called with ``anchors=None`` as in the
    reference, so a 'compress' DeformationNetwork fails here exactly like upstream), then ``decoder_shape``."""
    outs = []
    anchors = None
    with torch.no_grad():
        for points in torch.split(grid_points, nbatch_points, dim=1):
            n = points.shape[1]
            if encoding_expr is not None:
                e = encoding_expr.reshape(1, 1, -1)
                offsets, _ = decoder_expr(points, e.expand(1, n, e.shape[-1]), None)
                points_can = points + offsets
            else:
                points_can = points
            s = encoding_shape.reshape(1, 1, -1)
            out, anchors = decoder_shape(points_can, s.expand(1, n, s.shape[-1]), None)
            outs.append(out.reshape(-1).detach().cpu())
    logits = torch.cat(outs, dim=0).numpy()
    if return_anchors:
        return logits, anchors
    return logits


def deform_mesh(mesh, deformer, lat_rep, anchors, lat_rep_shape=None):
    """Forward-deform the vertices of a canonical mesh with ``deformer`` (one fused launch instead of the
    reference's 5000-vertex chunks + ``empty_cache``)."""
    points_neutral = torch.from_numpy(np.array(mesh.vertices)).float().unsqueeze(0).to(lat_rep.device)
    with torch.no_grad():
        cond = lat_rep if lat_rep_shape is None else torch.cat([lat_rep_shape, lat_rep], dim=-1)
        cond = cond.reshape(1, 1, -1)
        if anchors is not None:
            delta, _ = deformer(points_neutral, cond, anchors.reshape(1, -1, 3))
        else:
            delta, _ = deformer(points_neutral, cond.expand(1, points_neutral.shape[1], cond.shape[-1]), None)
    pred_posed = points_neutral[:, :, :3] + delta.reshape(1, -1, 3)
    verts = pred_posed.detach().cpu().squeeze(0).numpy()
    return make_mesh(verts, mesh.faces, process=False)

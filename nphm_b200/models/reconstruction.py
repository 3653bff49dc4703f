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


def _per_point(code: torch.Tensor, n: int) -> torch.Tensor:
    """latent of any shape -> ``(1, n, D)`` view (the layout the reference's modules are called with)."""
    code = code.reshape(1, 1, -1)
    return code.expand(1, n, code.shape[-1])


def _sweep(grid_points, chunk, evaluate):
    """Run ``evaluate(points) -> (values, anchors)`` over ``chunk``-sized pieces of the point axis; the values of every
    piece go to the host.  Returns (numpy (N,), anchors of the last piece)."""
    pieces, anchors = [], None
    with torch.no_grad():
        for start in range(0, grid_points.shape[1], chunk):
            values, anchors = evaluate(grid_points[:, start:start + chunk])
            pieces.append(values.reshape(-1).detach().cpu())
    return torch.cat(pieces).numpy(), anchors


def get_logits(decoder, encoding, grid_points, nbatch_points=100000, return_anchors=False):
    """SDF of ``grid_points`` (1 x N x 3) under latent ``encoding`` -> numpy (N,) float32 [, anchors]."""
    fused = (isinstance(decoder, FastEnsembleDeepSDFMirrored) and grid_points.is_cuda
             and grid_points.dtype == torch.float32)
    if fused:
        # one launch over all points; the chunk size only matters through the eval-mode last-point-of-chunk quirk
        with torch.no_grad():
            pts = grid_points.reshape(1, -1, 3)
            sdf, anchors = decoder.engine().query(pts, _as_row(encoding).to(pts.device), eval_quirk=not decoder.training,
                                                  quirk_period=None if decoder.training else int(nbatch_points))
            logits = _to_host(sdf.reshape(-1))
    else:
        logits, anchors = _sweep(grid_points, nbatch_points,
                                 lambda pts: decoder(pts, _per_point(encoding, pts.shape[1]), None))
    return (logits, anchors) if return_anchors else logits


def get_logits_backward(decoder_shape, decoder_expr, encoding_shape, encoding_expr, grid_points,
                        nbatch_points=100000, return_anchors=False):
    """Backward-warp query: points are first offset by ``decoder_expr`` (called with ``anchors=None`` as in the
    reference, so a 'compress' DeformationNetwork fails here exactly like upstream), then ``decoder_shape``."""
    def evaluate(pts):
        if encoding_expr is not None:
            pts = pts + decoder_expr(pts, _per_point(encoding_expr, pts.shape[1]), None)[0]
        return decoder_shape(pts, _per_point(encoding_shape, pts.shape[1]), None)

    logits, anchors = _sweep(grid_points, nbatch_points, evaluate)
    return (logits, anchors) if return_anchors else logits


def deform_mesh(mesh, deformer, lat_rep, anchors, lat_rep_shape=None):
    """Forward-deform the vertices of a canonical mesh with ``deformer`` (one fused launch instead of the
    reference's 5000-vertex chunks + ``empty_cache``)."""
    cached = getattr(mesh, '_nphm_device_vertices', None)
    if (cached is not None and cached.device == lat_rep.device and cached.shape[0] == len(mesh.vertices)):
        rest = cached[None]                      # vertices never left the device since marching cubes (mesh_from_logits)
    else:
        rest = torch.as_tensor(np.asarray(mesh.vertices), dtype=torch.float32, device=lat_rep.device)[None]
    code = lat_rep if lat_rep_shape is None else torch.cat([lat_rep_shape, lat_rep], dim=-1)
    with torch.no_grad():
        if anchors is None:
            shift = deformer(rest, _per_point(code, rest.shape[1]), None)[0]
        else:
            shift = deformer(rest, code.reshape(1, 1, -1), anchors.reshape(1, -1, 3))[0]
    posed = (rest[..., :3] + shift.reshape(1, -1, 3))[0]
    return make_mesh(posed.cpu().numpy(), mesh.faces, process=False)

"""PyTorch-composite implementations shared by the mirrored modules (autograd / CPU path).

The fused kernels are the product path; these functions compute the same maps with stock torch ops so that training
(double backward for the eikonal / normal losses), CPU tests and odd configurations keep working.  They are written
against the maths in SURVEY.md section 8a', not against the reference's code, and are pinned to the reference's outputs by
the golden fixtures (tests/test_host.py)."""
from __future__ import annotations

import math
from typing import Sequence

import torch

SQRT2 = math.sqrt(2.0)


def ensembled_affine(x, weight, bias, set_of_member):
    """x: A x M x D_in; weight: S x D_out x D_in; member a uses weight set ``set_of_member[a]``  ->  A x M x D_out."""
    w = weight.index_select(0, set_of_member)
    y = torch.einsum('amk,ank->amn', x, w)
    if bias is not None:
        y = y + bias.index_select(0, set_of_member)[:, None, :]
    return y


def skip_mlp(inp, layers: Sequence, skip_in, activation):
    """DeepSDF-style stack: ``layers[i]`` are callables, the network input is re-injected (concatenated, the sum scaled by
    1/sqrt(2)) in front of every layer listed in ``skip_in``; ``activation`` after all layers but the last."""
    h = inp
    for i, layer in enumerate(layers):
        if i in skip_in:
            h = torch.cat([h, inp], dim=-1) / SQRT2
        h = layer(h)
        if i + 1 < len(layers):
            h = activation(h)
    return h


def gaussian_blend(queries, centres, features, var, background):
    """Blend per-centre features with weights exp(-(|c - q| + 1e-5)^2 / var) (plus a constant -0.2/var logit for an extra
    background feature), normalised by sum + 1e-6.  queries B x N x 3, centres B x K x 3, features B x N x K(+1) x C."""
    gap = (centres[:, None, :, :] - queries[:, :, None, :]).norm(dim=-1) + 10e-6
    logits = -(gap * gap)
    if background:
        logits = torch.cat([logits, logits.new_full(logits.shape[:2] + (1,), -0.2)], dim=-1)
    w = torch.exp(logits / var)
    w = w / (w.sum(dim=-1, keepdim=True) + 1e-6)
    return (features * w[..., None]).sum(dim=2)


def ensemble_sdf(module, xyz, lat_rep):
    """Composite forward of ``FastEnsembleDeepSDFMirrored``: anchors from the global code, member-local (mirrored)
    coordinates, per-member condition [z_glob | z_k], member MLPs, Gaussian blend.  Returns (sdf B x N x 1, anchors)."""
    B, N, _ = xyz.shape
    K, G, L = module.num_kps, module.lat_dim_glob, module.lat_dim_loc
    if lat_rep.shape[1] == 1:
        lat_rep = lat_rep.expand(B, N, module.lat_dim)
    anchors = module.mlp_pos(lat_rep[:, 0, :G]).view(B, K, 3) + module.mean_anchors(xyz.device, xyz.dtype)[None]

    # member frames: anchor-centred for the K local members, the world frame for the last one; odd members of the
    # symmetric pairs look at the x-mirrored point
    origin = torch.cat([anchors, anchors.new_zeros(B, 1, 3)], dim=1)
    sign = xyz.new_ones(K + 1, 3)
    sign[1:2 * module.num_symm_pairs:2, 0] = -1.0
    local = (xyz[:, :, None, :] - origin[:, None, :, :]) * sign
    cond = torch.cat([lat_rep[:, :, None, :G].expand(B, N, K + 1, G), lat_rep[:, :, G:].reshape(B, N, K + 1, L)], dim=-1)

    s = module.ensembled_deep_sdf(local.permute(2, 0, 1, 3), cond.permute(2, 0, 1, 3))      # members x B x N x 1
    if not module.training:
        # the reference's eval-mode hack (EnsembledDeepSDF.py:260-261) hits the POINT axis: last point of the call -> 1
        s = s.clone()
        s[:, :, -1, 0] = 1
    return gaussian_blend(xyz[..., :3], anchors, s.permute(1, 2, 0, 3), var=0.1 ** 2, background=True), anchors

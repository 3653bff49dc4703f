"""Drop-in mirror of the reference module ``NPHM.models.iterative_root_finding`` (SNARF/DEQ-style Broyden search).

  * ``broyden``  src/NPHM/models/iterative_root_finding.py:5-71
  * ``nabla``    src/NPHM/models/iterative_root_finding.py:75-87
  * ``search``   src/NPHM/models/iterative_root_finding.py:91-168

Same signatures, thresholds (cvg 1e-6, dvg 0.2, 15 steps in ``search``) and return dictionaries.  The network
evaluations inside ``broyden`` happen under ``no_grad`` and therefore run on the fused MLP kernel.
"""
import torch

from .diff_operators import gradient, jac


def broyden(g, x_init, J_inv_init, max_steps=50, cvg_thresh=1e-5, dvg_thresh=1, eps=1e-6):
    """Batched Broyden root finder for g(x) = 0 (x: N x D x 1, J_inv: N x D x D), same update rule, freeze logic and
    return dictionary as the reference.  Written in the dense, flag-based form of the device kernel (`csrc/broyden.cu`):
    frozen samples receive a zero step instead of being gathered out, so ``g`` is always asked for all rows (it gets an
    all-true ``mask``); rows of ``g`` must therefore be independent of each other, which holds for `search`.

    Reference quirk kept: its ``x_opt`` aliases ``x``, so 'result' is the point where a sample stopped moving."""
    x = x_init.detach().clone()
    J = J_inv_init.detach().clone()
    n = x.shape[0]
    everyone = torch.ones(n, dtype=torch.bool, device=x.device)
    zero = torch.zeros((), dtype=x.dtype, device=x.device)
    res = g(x, mask=everyone)
    step = -torch.matmul(J, res)
    smallest = res.squeeze(-1).norm(dim=-1)
    live = everyone.clone()
    for _ in range(max_steps):
        sel = live.view(n, 1, 1)
        dx = torch.where(sel, step, zero)
        x = x + dx
        dres = torch.where(sel, g(x, mask=everyone) - res, zero)
        res = res + dres
        size = res.squeeze(-1).norm(dim=-1)
        smallest = torch.where(size < smallest, size, smallest)
        live = (smallest > cvg_thresh) & (size < dvg_thresh)
        if not bool(live.any()):
            break
        row = torch.matmul(dx.transpose(1, 2), J)                       # N x 1 x D
        col = dx - torch.matmul(J, dres)                                # N x D x 1
        denom = torch.matmul(row, dres)                                 # N x 1 x 1
        denom = denom + torch.where(denom >= 0, eps, -eps)
        J = torch.where(live.view(n, 1, 1), J + torch.matmul(col / denom, row), J)
        step = -torch.matmul(J, res)
    return {'result': x, 'diff': smallest, 'valid_ids': smallest < cvg_thresh}


def nabla(decoder_shape, xc, cond, anchors):
    """SDF and its spatial gradient at ``xc`` (sets requires_grad on xc like the reference)."""
    xc.requires_grad_(True)
    sdf_pred, _ = decoder_shape(xc, cond, anchors)
    return sdf_pred, gradient(sdf_pred, xc)


def _fused_search_condition(decoder_expr, xc, cond, anchors):
    """Per-query condition rows (B x C) when the device-side Broyden search applies: a `DeformationNetwork` in eval mode
    on CUDA whose condition is constant per query (modes compress / expr_only / glob_only), else None."""
    from .deepSDF import DeformationNetwork
    from .. import _native
    if not isinstance(decoder_expr, DeformationNetwork) or decoder_expr.training:
        return None
    if decoder_expr.mode not in ('compress', 'expr_only', 'glob_only') or not xc.is_cuda:
        return None
    with torch.no_grad():
        if not decoder_expr.defDeepSDF._fused_ok(xc.detach(), cond.detach()):
            return None
        lat = _native.constant_latent_rows(cond.detach())
        if lat is None:
            return None
        return decoder_expr._condition(xc.detach(), lat.unsqueeze(1), anchors.detach() if anchors is not None else None,
                                       per_point=False)[:, 0].contiguous()


def search(obs, cond, decoder_expr, anchors, multi_corresp=True):
    """Canonical correspondences of observed (posed) points: roots of x + F_ex(x) - obs.
    obs: B x N x 3.  Returns (xc_opt, result dict with 'valid_ids')."""
    B, N, _ = obs.shape
    starts = 5 if multi_corresp else 1
    target = obs
    guess = obs.detach().clone()
    if multi_corresp:
        # five starts per point: the point itself and four Gaussian perturbations (sigma 0.05), as the reference draws them
        jitter = 0.05 * torch.randn(B, N, starts, 3, device=obs.device)
        jitter[:, :, 0] = 0
        guess = (guess[:, :, None, :] + jitter).reshape(B, N * starts, 3)
        target = obs.repeat_interleave(starts, dim=1)
        cond = cond[:, :1].expand(B, N * starts, cond.shape[-1]).contiguous()
        if anchors is not None:
            anchors = anchors[:, :1].expand(B, N * starts, *anchors.shape[2:]).contiguous()
    M = guess.shape[1]

    J0_inv = jac(decoder_expr, guess, cond, anchors).inverse().reshape(B * M, 3, 3)

    def residual(flat_x, mask=None):
        pts = flat_x.reshape(B, M, 3)
        moved = decoder_expr(pts, cond, anchors)[0] + pts
        out = (moved - target).reshape(B * M, 3, 1)
        return out if mask is None else out[mask]

    per_query = _fused_search_condition(decoder_expr, guess, cond, anchors)
    with torch.no_grad():
        if per_query is not None:
            # whole iteration on the device: one fused-MLP launch + one 3x3 update kernel per step (nphm_mlp_broyden_search)
            x, diff, ok, _ = decoder_expr.defDeepSDF.engine().broyden_search(
                target, per_query, guess, J0_inv.reshape(B, M, 3, 3), max_steps=15, cvg_thresh=1e-6, dvg_thresh=0.2)
            result = {'result': x.reshape(B * M, 3, 1), 'diff': diff.reshape(-1), 'valid_ids': ok.reshape(-1)}
        else:
            result = broyden(residual, guess.reshape(B * M, 3, 1), J0_inv, cvg_thresh=1e-6, dvg_thresh=0.2, max_steps=15)

    shape = (B, N, starts) if multi_corresp else (B, N)
    result['valid_ids'] = result['valid_ids'].reshape(shape)
    return result['result'].reshape(*shape, 3), result

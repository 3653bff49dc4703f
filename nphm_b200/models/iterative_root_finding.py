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
    """Batched Broyden root finder for g(x) = 0.  x: N x D x 1, J_inv: N x D x D.
    Returns {'result': best x, 'diff': |g| at the best x, 'valid_ids': converged mask}."""
    x = x_init.clone().detach()
    J_inv = J_inv_init.clone().detach()
    active = torch.ones(x.shape[0], dtype=torch.bool, device=x.device)
    gx = g(x, mask=active)
    update = -J_inv.bmm(gx)
    x_opt = x
    best = torch.linalg.norm(gx.squeeze(-1), dim=-1)
    delta_gx = torch.zeros_like(gx)
    delta_x = torch.zeros_like(x)
    active = torch.ones_like(best, dtype=torch.bool)
    for _ in range(max_steps):
        delta_x[active] = update
        x[active] += delta_x[active]
        delta_gx[active] = g(x, mask=active) - gx[active]
        gx[active] += delta_gx[active]
        norm = torch.linalg.norm(gx.squeeze(-1), dim=-1)
        better = norm < best
        best[better] = norm.clone().detach()[better]
        x_opt[better] = x.clone().detach()[better]
        active = (best > cvg_thresh) & (norm < dvg_thresh)
        if active.sum() <= 0:
            break
        vT = delta_x[active].transpose(-1, -2).bmm(J_inv[active])
        a = delta_x[active] - J_inv[active].bmm(delta_gx[active])
        b = vT.bmm(delta_gx[active])
        b[b >= 0] += eps
        b[b < 0] -= eps
        J_inv[active] += (a / b).bmm(vT)
        update = -J_inv[active].bmm(gx[active])
    return {'result': x_opt, 'diff': best, 'valid_ids': best < cvg_thresh}


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
    n_batch, n_point, _ = obs.shape
    if multi_corresp:
        num_inits = 5
        xc_init = obs.detach().clone().unsqueeze(2).repeat(1, 1, num_inits, 1)
        offsets = torch.randn(xc_init.shape, device=xc_init.device) * 0.05
        offsets[:, :, 0, :] = 0
        xc_init = (xc_init + offsets).reshape(n_batch, n_point * num_inits, 3)
        obs = obs.repeat_interleave(num_inits, dim=1)
        cond = cond[:, 0, :].unsqueeze(1).repeat(1, xc_init.shape[1], 1)
        if anchors is not None:
            anchors = anchors[:, 0, :, :].unsqueeze(1).repeat(1, xc_init.shape[1], 1, 1)
    else:
        xc_init = obs.detach().clone()

    J_inv_init = jac(decoder_expr, xc_init, cond, anchors).inverse()
    xc_init = xc_init.reshape(-1, 3, 1)
    J_inv_init = J_inv_init.flatten(0, 1)

    def residual(xc_opt, mask=None):
        pts = xc_opt.reshape(n_batch, -1, 3)
        off, _ = decoder_expr(pts, cond, anchors)
        err = (off + pts) - obs
        return err.flatten(0, 1)[mask].unsqueeze(-1)

    fused_cond = _fused_search_condition(decoder_expr, xc_init.reshape(n_batch, -1, 3), cond, anchors)
    with torch.no_grad():
        if fused_cond is not None:
            # whole iteration on the device: one fused-MLP launch + one 3x3 update kernel per step (nphm_mlp_broyden_search)
            x, diff, valid, _ = decoder_expr.defDeepSDF.engine().broyden_search(
                obs, fused_cond, xc_init.reshape(n_batch, -1, 3), J_inv_init.reshape(n_batch, -1, 3, 3),
                max_steps=15, cvg_thresh=1e-6, dvg_thresh=0.2)
            result = {'result': x.reshape(-1, 3, 1), 'diff': diff.reshape(-1), 'valid_ids': valid.reshape(-1)}
        else:
            result = broyden(residual, xc_init, J_inv_init, cvg_thresh=1e-6, dvg_thresh=0.2, max_steps=15)

    if multi_corresp:
        xc_opt = result['result'].reshape(n_batch, n_point, -1, 3)
        result['valid_ids'] = result['valid_ids'].reshape(n_batch, n_point, num_inits)
    else:
        xc_opt = result['result'].reshape(n_batch, n_point, 3)
        result['valid_ids'] = result['valid_ids'].reshape(n_batch, n_point)
    return xc_opt, result

"""Drop-in mirror of the reference module ``NPHM.models.fitting``.

  * ``inference_identity_space``                 src/NPHM/models/fitting.py:180-285
  * ``inference_iterative_root_finding_joint``   src/NPHM/models/fitting.py:14-177

``inference_identity_space`` runs on the fused fitting kernels (``nphm_fit_identity_step``: forward, clamped |sdf|
loss, analytic latent gradient and Adam in five launches, no autograd graph) when the decoder is a
``FastEnsembleDeepSDFMirrored`` on a CUDA device; otherwise it falls back to autograd through the composite modules,
written here independently but following the same loop.  Host-side behaviour kept from the reference: the
point sampling draws from torch's global CPU generator in the same order (``torch.randint`` at :214,:219), the
``lambdas`` dict is mutated by the schedule (:199-206), Adam runs with lr 0.01*lr_scale halved by the schedule, and the
returned anchors are those of the LAST iteration's pre-update latent (:211).
"""
from __future__ import annotations

import ctypes
import os
from typing import Dict, List

import torch
from torch import optim

from .. import _native
from .EnsembledDeepSDF import FastEnsembleDeepSDFMirrored
from .iterative_root_finding import search, nabla          # noqa: F401  (same names as the reference imports)
from .diff_operators import gradient, jac                    # noqa: F401

NUM_OBSERVATIONS_PER_BATCH = 5
NUM_POINTS_PER_OBSERVATION = 1000


def _apply_schedule(j, step_scale, schedule_cfg, lambdas, lr):
    """Eye-balled schedule of the reference (:199-206 / :40-52): divides lr and lambdas at fixed iterations."""
    key = int(j / step_scale)
    if key in schedule_cfg.get('lr', {}):
        lr = lr / schedule_cfg['lr'][key]
    for cfg_key, lam_key in (('symm_dist', 'symm_dist'), ('reg_glob', 'reg_global'), ('reg_loc', 'reg_loc'),
                             ('reg_expr', 'reg_expr')):
        if cfg_key in schedule_cfg and key in schedule_cfg[cfg_key] and (lam_key in lambdas or cfg_key != 'reg_expr'):
            lambdas[lam_key] /= schedule_cfg[cfg_key][key]
    return lr


def _clamp_for_iteration(j, step_scale):
    """Sequential strict filters l<0.1, (j>250) l<0.05, (j>500) l<0.0075 collapse to one threshold (:239-246)."""
    clamp = 0.1
    if j > int(250 * step_scale):
        clamp = 0.05
    if j > int(500 * step_scale):
        clamp = 0.0075
    return clamp


def _sample_observations(all_obs):
    """Reference sampling order on the global CPU generator (:214-222)."""
    num_observations = len(all_obs)
    sampled_observations_idx = torch.randint(0, num_observations, [NUM_OBSERVATIONS_PER_BATCH])
    sampled_points = []
    for i in range(NUM_OBSERVATIONS_PER_BATCH):
        sampled_idx = sampled_observations_idx[i]
        n_samps = min(NUM_POINTS_PER_OBSERVATION, all_obs[sampled_idx].shape[0])
        subsample_idx = torch.randint(0, all_obs[sampled_idx].shape[0], [n_samps])
        sampled_points.append(all_obs[sampled_idx][subsample_idx.to(all_obs[sampled_idx].device), :])
    return torch.stack(sampled_points, dim=0), sampled_observations_idx


def _current_anchors(decoder, lat_rep_shape, device):
    """Anchors of the current identity code, with graph.  The reference evaluates the whole decoder on a dummy point for
    this (fitting.py:57, :218); the drop-in ensemble exposes the anchor head directly (same values, ~50 launches less)."""
    if isinstance(decoder, FastEnsembleDeepSDFMirrored):
        return decoder.predict_anchors(lat_rep_shape)
    return decoder(torch.zeros([1, 1, 3], device=device), lat_rep_shape, None)[1]


def _fused_identity(decoder) -> bool:
    """The fused step implements the training-mode forward, which is how the reference runs its fitters
    (scripts/fitting/fitting_pointclouds.py:268 calls ``decoder_shape.train()`` first).  In eval mode the reference's
    forward overwrites the last point of every row (EnsembledDeepSDF.py:257-259): that case takes the autograd path."""
    p = next(decoder.parameters())
    return (isinstance(decoder, FastEnsembleDeepSDFMirrored) and p.is_cuda and decoder.training
            and decoder.ensembled_deep_sdf.num_layers == 6)


class IdentityFitter:
    """Stateful wrapper around ``nphm_fit_identity_step``: latent + Adam moments live on the device."""

    def __init__(self, decoder: FastEnsembleDeepSDFMirrored, device):
        self.decoder = decoder
        self.device = device
        self.engine = decoder.engine()
        self.latent = torch.zeros(decoder.lat_dim, device=device, dtype=torch.float32)
        self.m = torch.zeros_like(self.latent)
        self.v = torch.zeros_like(self.latent)
        self.loss_terms = torch.zeros(8, device=device, dtype=torch.float32)
        self.grad = torch.zeros_like(self.latent)
        self.t = 0

    def step(self, points: torch.Tensor, lambdas: Dict[str, float], clamp: float, lr: float, apply_update: bool = True):
        pts = points.reshape(-1, 3).to(dtype=torch.float32).contiguous()
        if apply_update:
            self.t += 1
        fp = _native.FitParams(float(lambdas.get('surface', 0.0)), float(lambdas.get('reg_global', 0.0)),
                               float(lambdas.get('reg_loc', 0.0)), float(lambdas.get('reg_unobserved', 0.0)),
                               float(lambdas.get('symm_dist', 0.0)), float(clamp), float(lr), max(self.t, 1))
        with torch.cuda.device(self.device):
            _native.check(_native.lib().nphm_fit_identity_step(
                self.engine.handle, pts.data_ptr(), pts.shape[0], self.latent.data_ptr(), self.m.data_ptr(),
                self.v.data_ptr(), ctypes.byref(fp), int(apply_update), self.loss_terms.data_ptr(),
                self.grad.data_ptr(), None, torch.cuda.current_stream(self.device).cuda_stream),
                'nphm_fit_identity_step')


class _FusedSurfaceLoss(torch.autograd.Function):
    """``sdf = decoder(xc, z_id); sdf[valid].abs()[< clamp].mean()`` of the joint fitter (reference fitting.py:114-125) as
    one native call (`nphm_fit_surface_grad`): tensor-core forward, analytic backward w.r.t. the identity code (member
    inputs, anchors, blend weights) and w.r.t. the query points.  Autograd continues from the point gradient into the
    implicit-differentiation correction and the deformation network."""

    @staticmethod
    def forward(ctx, xc, lat_rep_shape, valid, clamp, decoder):
        dev = xc.device
        pts = xc.detach().reshape(-1, 3).to(torch.float32).contiguous()
        lat = lat_rep_shape.detach().reshape(-1).to(torch.float32).contiguous()
        mask = valid.reshape(-1).to(torch.uint8).contiguous()
        eng = decoder.engine()
        terms = torch.empty(8, device=dev, dtype=torch.float32)
        g_lat = torch.empty_like(lat)
        g_pts = torch.empty_like(pts)
        with torch.cuda.device(dev):
            _native.check(_native.lib().nphm_fit_surface_grad(
                eng.handle, pts.data_ptr(), pts.shape[0], lat.data_ptr(), mask.data_ptr(), float(clamp), terms.data_ptr(),
                g_lat.data_ptr(), g_pts.data_ptr(), None, torch.cuda.current_stream(dev).cuda_stream),
                'nphm_fit_surface_grad')
        ctx.save_for_backward(g_lat, g_pts)
        ctx.shapes = (xc.shape, lat_rep_shape.shape)
        return terms[0].clone()

    @staticmethod
    def backward(ctx, grad_out):
        g_lat, g_pts = ctx.saved_tensors
        return grad_out * g_pts.reshape(ctx.shapes[0]), grad_out * g_lat.reshape(ctx.shapes[1]), None, None, None


def inference_identity_space(decoder,
                             all_obs: List[torch.Tensor],
                             lambdas,
                             n_steps,
                             schedule_cfg: Dict,
                             step_scale=1,
                             lr_scale=1):
    """Fit the identity code to point-cloud observations.  Returns ``(lat_rep_shape (1,1,D), anchors (1,K,3))``."""
    device = all_obs[0].device
    lr = 0.01 * lr_scale
    if _fused_identity(decoder) and device.type == 'cuda':
        fitter = IdentityFitter(decoder, device)
        z_prev = fitter.latent.clone()
        for j in range(int(n_steps * step_scale)):
            lr = _apply_schedule(j, step_scale, schedule_cfg, lambdas, lr)
            obs, _ = _sample_observations(all_obs)
            z_prev.copy_(fitter.latent)
            fitter.step(obs, lambdas, _clamp_for_iteration(j, step_scale), lr)
        with torch.no_grad():
            _, anchors = fitter.engine.query(torch.zeros(1, 1, 3, device=device), z_prev.reshape(1, -1), eval_quirk=False)
        lat_rep_shape = fitter.latent.reshape(1, 1, -1).clone().requires_grad_(True)
        return lat_rep_shape, anchors

    # ---- autograd path (any decoder, CPU or GPU)
    lat_dim = decoder.lat_dim
    lat_rep_shape = torch.zeros([1, 1, lat_dim], device=device, requires_grad=True)
    opt = optim.Adam(params=[lat_rep_shape], lr=lr)
    anchors = None
    for j in range(int(n_steps * step_scale)):
        new_lr = _apply_schedule(j, step_scale, schedule_cfg, lambdas, lr)
        if new_lr != lr:
            lr = new_lr
            for group in opt.param_groups:
                group['lr'] = lr
        opt.zero_grad()
        anchors = _current_anchors(decoder, lat_rep_shape, device)
        obs, _ = _sample_observations(all_obs)
        nb = obs.shape[0]
        if hasattr(decoder, 'lat_dim_loc'):
            sdf, _ = decoder(obs, lat_rep_shape.repeat(nb, 1, 1), None)
        else:
            sdf, _ = decoder(obs, lat_rep_shape.repeat(nb, obs.shape[1], 1), None)
        l = sdf.abs()
        l = l[l < _clamp_for_iteration(j, step_scale)]
        loss_dict = {'surface': l.mean()}
        loss_dict.update(_latent_regularisers(decoder, lat_rep_shape))
        loss = 0
        for k in lambdas.keys():
            loss = loss + loss_dict[k] * lambdas[k]
        loss.backward()
        opt.step()
    return lat_rep_shape, anchors


def _latent_regularisers(decoder, lat_rep_shape):
    """reg_loc / reg_global / reg_unobserved / symm_dist of the reference (:252-272)."""
    out = {}
    if hasattr(decoder, 'lat_dim_glob'):
        g, lc = decoder.lat_dim_glob, decoder.lat_dim_loc
        out['reg_loc'] = (torch.norm(lat_rep_shape[..., g:], dim=-1) ** 2).mean()
        out['reg_global'] = (torch.norm(lat_rep_shape[..., :g], dim=-1) ** 2).mean()
        ru = 0
        for idx in (30, 31, 39):
            ru = ru + torch.norm(lat_rep_shape[..., g + idx * lc:g + (idx + 1) * lc], dim=-1).square().mean()
        out['reg_unobserved'] = ru
        n_pairs = decoder.num_symm_pairs
        loc = lat_rep_shape[:, :, g:g + 2 * n_pairs * lc].view(lat_rep_shape.shape[0], n_pairs * 2, lc)
        out['symm_dist'] = torch.norm(loc[:, ::2, :] - loc[:, 1::2, :], dim=-1).mean()
    else:
        out['symm_dist'] = 0
        out['reg_unobserved'] = 0
        out['reg_loc'] = 0
        out['reg_global'] = (torch.norm(lat_rep_shape, dim=-1) ** 2).mean()
    return out


def _native_joint(decoder, decoder_expr, device) -> bool:
    """The autograd-free joint fitter applies to the shipped configuration: fused identity ensemble (training mode) and a
    'compress' DeformationNetwork in eval mode on the same CUDA device."""
    from .deepSDF import DeformationNetwork
    if device.type != 'cuda' or not _fused_identity(decoder) or not isinstance(decoder_expr, DeformationNetwork):
        return False
    if decoder_expr.training or decoder_expr.mode != 'compress':
        return False
    bb = decoder_expr.defDeepSDF
    return next(bb.parameters()).is_cuda and bb.num_freq_bands is None and bb.beta == 100 and bb.out_dim_net == 3


class JointFitter:
    """One iteration of ``inference_iterative_root_finding_joint`` (reference fitting.py:38-175) WITHOUT an autograd graph.

    The reference builds xc = p - J^-1 (F_ex(p; theta) + p - stopgrad(.)) so that value(xc) = p and d xc / d theta =
    -J^-1 dF_ex/d theta (:99-106), then calls loss.backward() (:167).  Written out:
        g_x   = d surface / d xc                                   nphm_fit_surface_grad (with d surface / d z_id)
        u     = -J^-T g_x            per point, J = I + dF_ex/dx    nphm_mlp_inverse_jacobian (forward-mode, tensor cores)
        g_c   = sum_n (dF_ex/d cond)^T u_n   per sampled scan       nphm_mlp_backward_inputs (adjoint pass, tensor cores)
        cond  = [compressor([z_id | anchors]) (32) | z_ex (200)]    -> z_ex rows, and through the compressor to z_id and the
                                                                      anchors (mlp_pos backward inside nphm_fit_apply_gradient)
    followed by the regularisers and the two Adam updates (nphm_fit_apply_gradient, nphm_adam_step)."""

    def __init__(self, decoder, decoder_expr, num_observations: int, device):
        self.dec, self.dfn, self.device = decoder, decoder_expr, device
        self.eng = decoder.engine()
        self.mlp = decoder_expr.defDeepSDF.engine()
        E = decoder_expr.lat_dim_expr
        self.z_id = torch.zeros(decoder.lat_dim, device=device)
        self.m_id, self.v_id = torch.zeros_like(self.z_id), torch.zeros_like(self.z_id)
        self.z_ex = torch.zeros(num_observations, E, device=device)
        self.m_ex, self.v_ex = torch.zeros_like(self.z_ex), torch.zeros_like(self.z_ex)
        self.terms = torch.zeros(8, device=device)
        self.loss_terms = torch.zeros(8, device=device)
        self.t = 0
        self.anchors = None
        # Broyden early exit (the reference leaves its loop when nobody is active) costs a host sync every third step; without it
        # an iteration is a pure launch sequence.  Default: on for small batches is not worth a sync on a B200 - off.
        self.early_exit = bool(int(os.environ.get('NPHM_BROYDEN_EARLY_EXIT', '0')))

    def step(self, obs, obs_idx, lambdas, clamp, lr, apply_update: bool = True):
        """One iteration; with ``apply_update=False`` nothing is modified and ``(d loss / d z_id, d loss / d z_ex)`` - the
        tensors the reference hands to its two ``Adam.step()`` calls - are returned."""
        nat, dev = _native, self.device
        nb, n_point, _ = obs.shape
        E, D = self.z_ex.shape[1], self.z_id.shape[0]
        if apply_update:
            self.t += 1
        stream = torch.cuda.current_stream(dev).cuda_stream
        with torch.no_grad(), torch.cuda.device(dev):
            anchors = self.eng.anchors(self.z_id[None])                                  # 1 x K x 3, pre-update code (:59)
            self.anchors = anchors
            # condition of the deformation backbone (deepSDF.py:218-219): compressor([z_id | anchors]) (32) | z_ex (E)
            lin = self.dfn.compressor[0] if isinstance(self.dfn.compressor, torch.nn.Sequential) else self.dfn.compressor
            Wc, bc = lin.weight, lin.bias
            first = torch.cat([self.z_id, anchors.reshape(-1)])
            c32 = (Wc * first[None, :]).sum(1) + bc
            cond = torch.cat([c32[None, :].expand(nb, -1), self.z_ex[obs_idx]], dim=1).contiguous()
            # correspondence search (iterative_root_finding.py:91-168): J0^-1 at the observed points, Broyden on the device
            obs = obs.to(torch.float32).contiguous()
            _, j0_inv = self.mlp.inverse_jacobian(obs, cond)
            p, _, valid, _ = self.mlp.broyden_search(obs, cond, obs, j0_inv, max_steps=15, cvg_thresh=1e-6, dvg_thresh=0.2,
                                                     early_exit=self.early_exit)
            # implicit differentiation of the root (:99-106): J^-1 at the root
            _, j_inv = self.mlp.inverse_jacobian(p, cond)
            # surface term (:111-136): value + d/d z_id + d/d xc
            pts = p.reshape(-1, 3)
            mask = valid.reshape(-1).to(torch.uint8).contiguous()
            g_lat = torch.empty(D, device=dev)
            g_pts = torch.empty_like(pts)
            nat.check(nat.lib().nphm_fit_surface_grad(self.eng.handle, pts.data_ptr(), pts.shape[0], self.z_id.data_ptr(),
                                                      mask.data_ptr(), float(clamp), self.terms.data_ptr(), g_lat.data_ptr(),
                                                      g_pts.data_ptr(), None, stream), 'nphm_fit_surface_grad')
            u = -(j_inv * g_pts.reshape(nb, n_point, 3, 1)).sum(-2)                       # -J^-T g_x
            g_cond, _ = self.mlp.backward_inputs(p, cond, u, reuse_value_pass=True)     # nb x (32 + E); same points as j_inv
            g_first = (Wc * g_cond[:, :32].sum(0)[:, None]).sum(0)                       # compressor^T -> [z_id | anchors]
            g_zid = (g_lat + g_first[:D]).contiguous()
            g_anchors = g_first[D:].contiguous()
            stats = torch.stack([self.terms[5], self.terms[0] * self.terms[5]]).nan_to_num_(0.0).contiguous()
            # expression codes: surface route + reg_expr = mean_b ||z_ex[idx_b]||^2 (:137), dense Adam over all rows (:169)
            lam_s, lam_e = float(lambdas.get('surface', 0.0)), float(lambdas.get('reg_expr', 0.0))
            g_zex = torch.zeros_like(self.z_ex)
            g_zex.index_add_(0, obs_idx, lam_s * g_cond[:, 32:] + (2.0 * lam_e / nb) * self.z_ex[obs_idx])
            fp = nat.FitParams(lam_s, float(lambdas.get('reg_global', 0.0)), float(lambdas.get('reg_loc', 0.0)),
                               float(lambdas.get('reg_unobserved', 0.0)), float(lambdas.get('symm_dist', 0.0)), float(clamp),
                               float(lr), max(self.t, 1))
            g_total = None if apply_update else torch.empty(D, device=dev)
            nat.check(nat.lib().nphm_fit_apply_gradient(self.eng.handle, self.z_id.data_ptr(), self.m_id.data_ptr(),
                                                        self.v_id.data_ptr(), ctypes.byref(fp), g_zid.data_ptr(), stats.data_ptr(),
                                                        g_anchors.data_ptr(), int(apply_update), self.loss_terms.data_ptr(),
                                                        None if apply_update else g_total.data_ptr(), stream),
                      'nphm_fit_apply_gradient')
            if not apply_update:
                return g_total, g_zex
            nat.check(nat.lib().nphm_adam_step(self.z_ex.data_ptr(), g_zex.data_ptr(), self.m_ex.data_ptr(), self.v_ex.data_ptr(),
                                               self.z_ex.numel(), float(lr), self.t, stream), 'nphm_adam_step')
            return None


def inference_iterative_root_finding_joint(decoder,
                                           decoder_expr,
                                           all_obs: List[torch.Tensor],
                                           lambdas,
                                           n_steps,
                                           schedule_cfg: Dict,
                                           step_scale=1,
                                           lr_scale=1):
    """Joint identity + expression fitting with Broyden correspondences (reference :14-177).

    Shipped configuration (fused ensemble + 'compress' DeformationNetwork on CUDA): no autograd graph at all, see
    :class:`JointFitter`.  Anything else: the correspondence search runs on the device (`nphm_mlp_broyden_search`), the
    surface term comes from `nphm_fit_surface_grad`, the deformation network's part of the chain stays on autograd.  Returns
    ``(lat_rep (n_obs,1,E), lat_rep_shape (1,1,D), anchors)``."""
    device = all_obs[0].device
    num_observations = len(all_obs)
    if _native_joint(decoder, decoder_expr, device) and not os.environ.get('NPHM_JOINT_AUTOGRAD'):
        fitter = JointFitter(decoder, decoder_expr, num_observations, device)
        lr = 0.01 * lr_scale
        for j in range(int(n_steps * step_scale)):
            lr = _apply_schedule(j, step_scale, schedule_cfg, lambdas, lr)
            obs, obs_idx = _sample_observations(all_obs)
            fitter.step(obs, obs_idx.long().to(device), lambdas, _clamp_for_iteration(j, step_scale), lr)
        lat_rep = fitter.z_ex.reshape(num_observations, 1, -1).clone().requires_grad_(True)
        lat_rep_shape = fitter.z_id.reshape(1, 1, -1).clone().requires_grad_(True)
        return lat_rep, lat_rep_shape, fitter.anchors
    lat_expr_dim = decoder_expr.lat_dim_expr if hasattr(decoder_expr, 'lat_dim_expr') else 200
    lat_rep = torch.zeros([num_observations, 1, lat_expr_dim], device=device, requires_grad=True)
    lat_rep_shape = torch.zeros([1, 1, decoder.lat_dim], device=device, requires_grad=True)
    lr = 0.01 * lr_scale
    opt = optim.Adam(params=[lat_rep_shape], lr=lr)
    opt_expr = optim.Adam(params=[lat_rep], lr=lr)
    anchors = None
    for j in range(int(n_steps * step_scale)):
        new_lr = _apply_schedule(j, step_scale, schedule_cfg, lambdas, lr)
        if new_lr != lr:
            lr = new_lr
            for o in (opt, opt_expr):
                for group in o.param_groups:
                    group['lr'] = lr
        opt.zero_grad()
        opt_expr.zero_grad()
        anchors = _current_anchors(decoder, lat_rep_shape, device)
        obs, obs_idx = _sample_observations(all_obs)
        obs_idx = obs_idx.long().to(device)
        nb, n_point, _ = obs.shape
        glob_cond = torch.cat([lat_rep_shape.repeat(nb, 1, 1), lat_rep[obs_idx, :, :]], dim=-1)
        has_local = hasattr(decoder, 'lat_dim_loc')
        search_anchors = anchors.clone().unsqueeze(1).repeat(nb, n_point, 1, 1) if has_local else None
        p_corresp, search_result = search(obs, glob_cond.repeat(1, n_point, 1), decoder_expr, search_anchors,
                                          multi_corresp=False)
        p_corresp = p_corresp.detach()
        _anchors = anchors.clone().unsqueeze(1).repeat(nb, n_point, 1, 1) if anchors is not None else None
        # implicit differentiation of the root: value(xc) = p, d xc = -J^-1 d F_ex
        preds_posed, _ = decoder_expr(p_corresp, glob_cond.repeat(1, n_point, 1), _anchors)
        preds_posed = preds_posed + p_corresp
        grad_inv = jac(decoder_expr, p_corresp, glob_cond.repeat(1, n_point, 1), _anchors).inverse()
        correction = preds_posed - preds_posed.detach()
        correction = torch.einsum('bnij,bnj->bni', -grad_inv.detach(), correction)
        xc = p_corresp + correction
        if has_local and _fused_identity(decoder) and xc.is_cuda:
            surface = _FusedSurfaceLoss.apply(xc, lat_rep_shape, search_result['valid_ids'],
                                              _clamp_for_iteration(j, step_scale), decoder)
        else:
            if has_local:
                sdf, _ = decoder(xc, lat_rep_shape.repeat(nb, 1, 1), None)
            else:
                sdf, _ = decoder(xc, lat_rep_shape.repeat(nb, n_point, 1), None)
            sdf = sdf[search_result['valid_ids'], :]
            l = sdf.abs()
            l = l[l < _clamp_for_iteration(j, step_scale)]
            surface = l.mean()
        loss_dict = {'surface': surface,
                     'reg_expr': (torch.norm(lat_rep[obs_idx, :, :], dim=-1) ** 2).mean()}
        loss_dict.update(_latent_regularisers(decoder, lat_rep_shape))
        loss = 0
        for k in lambdas.keys():
            loss = loss + loss_dict[k] * lambdas[k]
        loss.backward()
        opt.step()
        opt_expr.step()
    return lat_rep, lat_rep_shape, anchors

#!/usr/bin/env python
"""Generate the golden fixtures in this directory by running the UNMODIFIED reference.

Run in the build container only (the GPU box has no /root/reference):

    python tests/golden/make_golden.py

The reference (SimonGiebenhain/NPHM @ /root/reference) ships no tests and no golden vectors
(SURVEY.md section 4), so the pins are the outputs of its own PyTorch modules (torch 2.11, CPU, fp32) on
seeded random weights - default initialisation, ``torch.manual_seed`` - plus the three numeric assets
the hot path reads (assets/anchors_39.npy, assets/nphm_lat_{mean,std}.npy).  Weights are not stored:
the drop-in classes of ``nphm_b200.models`` reproduce the reference's initialisation bit for bit from
the same seed, which ``sha256`` fields below let the tests verify.

Third-party imports the reference makes at module top level but never uses on this path (trimesh,
mcubes, pytorch3d, pyvista) are stubbed in ``sys.modules``; ``NPHM.env_paths`` is never imported.
"""
import hashlib
import os
import sys
import types

import numpy as np
import torch

REF = '/root/reference'
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(REF, 'src'))
for name in ('trimesh', 'mcubes', 'pyvista', 'pytorch3d', 'pytorch3d.ops'):
    mod = types.ModuleType(name)
    sys.modules[name] = mod
sys.modules['pytorch3d.ops'].knn_points = None
sys.modules['pytorch3d.ops'].knn_gather = None
sys.modules['pytorch3d'].ops = sys.modules['pytorch3d.ops']

from NPHM.models.EnsembledDeepSDF import FastEnsembleDeepSDFMirrored     # noqa: E402
from NPHM.models.deepSDF import DeepSDF, DeformationNetwork              # noqa: E402
from NPHM.models.reconstruction import get_logits                        # noqa: E402
from NPHM.utils.reconstruction import create_grid_points_from_bounds     # noqa: E402
from NPHM.models.fitting import inference_identity_space, inference_iterative_root_finding_joint   # noqa: E402

MINI = [-.55, -.5, -.95]
MAXI = [0.55, 0.75, 0.4]


def sd_hash(sd):
    h = hashlib.sha256()
    for k in sorted(sd.keys()):
        h.update(k.encode())
        h.update(np.ascontiguousarray(sd[k].detach().cpu().numpy()).tobytes())
    return h.hexdigest()


def make_ensemble(seed, anchors, scale=1.0, pos_mlp_dim=256):
    torch.manual_seed(seed)
    dec = FastEnsembleDeepSDFMirrored(lat_dim_glob=64, lat_dim_loc=32, n_loc=39, n_symm_pairs=16,
                                      anchors=anchors, hidden_dim=200, n_layers=4, pos_mlp_dim=pos_mlp_dim)
    if scale != 1.0:
        with torch.no_grad():
            for i in range(5):
                getattr(dec.ensembled_deep_sdf, 'lin%d' % i).weight.mul_(scale)
    return dec


def sample_latent(seed, mean, std):
    torch.manual_seed(seed)
    return torch.randn(mean.shape) * std * 0.85 + mean        # fitting_pointclouds.py:206


def random_points(seed, n):
    rng = np.random.RandomState(seed)
    lo, hi = np.array(MINI), np.array(MAXI)
    return (rng.rand(n, 3) * (hi - lo) + lo).astype(np.float32)


def main():
    torch.set_num_threads(os.cpu_count())
    anchors64 = np.load(os.path.join(REF, 'assets', 'anchors_39.npy'))
    lat_mean = np.load(os.path.join(REF, 'assets', 'nphm_lat_mean.npy'))
    lat_std = np.load(os.path.join(REF, 'assets', 'nphm_lat_std.npy'))
    np.savez_compressed(os.path.join(HERE, 'assets.npz'), anchors_39=anchors64,
                        nphm_lat_mean=lat_mean, nphm_lat_std=lat_std)
    anchors = torch.from_numpy(anchors64).float().unsqueeze(0).unsqueeze(0)    # fitting_pointclouds.py:80

    # ---------------------------------------------------------------- grid helper
    grid5 = create_grid_points_from_bounds(MINI, MAXI, 5)
    grid16 = create_grid_points_from_bounds(MINI, MAXI, 16)

    # ---------------------------------------------------------------- ensemble forward
    out = {}
    for tag, seed, scale in (('a', 0, 1.0), ('b', 5, 2.0)):
        dec = make_ensemble(seed, anchors, scale)
        out['sha256_' + tag] = np.array(sd_hash(dec.state_dict()))
        lat = sample_latent(1 if tag == 'a' else 7, torch.from_numpy(lat_mean), torch.from_numpy(lat_std))
        if tag == 'b':
            lat = lat * 20.0          # large codes: exercises big activations / moving anchors
        pts = np.concatenate([random_points(2, 3000), grid16.astype(np.float32)], axis=0)
        x = torch.from_numpy(pts).unsqueeze(0)
        with torch.no_grad():
            dec.eval()
            sdf_eval, anc = dec(x.clone(), lat.reshape(1, 1, -1), None)
            dec.train()
            sdf_train, _ = dec(x.clone(), lat.reshape(1, 1, -1), None)
            dec.eval()
            grid20 = torch.from_numpy(create_grid_points_from_bounds(MINI, MAXI, 20)).float().unsqueeze(0)
            logits20 = get_logits(dec, lat, grid20, nbatch_points=3000)     # 8000 pts -> chunks 3000,3000,2000
        out['latent_' + tag] = lat.numpy()
        out['points_' + tag] = pts
        out['sdf_eval_' + tag] = sdf_eval.numpy().reshape(-1)
        out['sdf_train_' + tag] = sdf_train.numpy().reshape(-1)
        out['anchors_' + tag] = anc.numpy().reshape(39, 3)
        out['logits20_' + tag] = logits20.astype(np.float32)
        print('ensemble', tag, 'sdf range', float(sdf_train.min()), float(sdf_train.max()))
    out['grid5'] = grid5
    np.savez_compressed(os.path.join(HERE, 'ensemble.npz'), **out)

    # ---------------------------------------------------------------- deformation network + plain DeepSDF
    dec = make_ensemble(0, anchors)
    lat = sample_latent(1, torch.from_numpy(lat_mean), torch.from_numpy(lat_std))
    with torch.no_grad():
        _, anc = dec(torch.zeros(1, 1, 3), lat.reshape(1, 1, -1), None)
    torch.manual_seed(10)
    dfn = DeformationNetwork(mode='compress', lat_dim_expr=200, lat_dim_id=32, lat_dim_glob_shape=64,
                             lat_dim_loc_shape=32, n_loc=39, anchors=anchors, hidden_dim=512, nlayers=6,
                             out_dim=3, input_dim=3)
    dfn.eval()
    torch.manual_seed(11)
    z_ex = torch.randn(200) * 0.1
    pts = random_points(3, 2000)
    with torch.no_grad():
        cond = torch.cat([lat, z_ex]).reshape(1, 1, -1).repeat(1, pts.shape[0], 1)
        off, last = dfn(torch.from_numpy(pts).unsqueeze(0), cond, anc)
    torch.manual_seed(12)
    npm = DeepSDF(lat_dim=64, hidden_dim=96, nlayers=8, geometric_init=True)      # small NPM-style SDF
    torch.manual_seed(13)
    z_npm = torch.randn(64) * 0.1
    with torch.no_grad():
        npm_out, _ = npm(torch.from_numpy(pts).unsqueeze(0), z_npm.reshape(1, 1, -1).repeat(1, pts.shape[0], 1))
    np.savez_compressed(os.path.join(HERE, 'deform.npz'),
                        sha256_def=np.array(sd_hash(dfn.state_dict())), sha256_npm=np.array(sd_hash(npm.state_dict())),
                        latent_id=lat.numpy(), z_ex=z_ex.numpy(), anchors=anc.numpy().reshape(39, 3), points=pts,
                        offsets=off.numpy().reshape(-1, 3), last=last.numpy().reshape(-1),
                        z_npm=z_npm.numpy(), npm_out=npm_out.numpy().reshape(-1))
    print('deform offsets range', float(off.min()), float(off.max()), 'npm', float(npm_out.min()), float(npm_out.max()))

    # ---------------------------------------------------------------- identity-space fitting trajectory
    # 3 synthetic observations of 300 points; schedule compressed with step_scale=0.01 so that 10 iterations
    # pass every lr / lambda / clamp event of fitting_pointclouds.py:253-266.  The reference code runs
    # unmodified; torch.optim.Adam is wrapped only to RECORD the gradient / parameter it is handed each step.
    import torch.optim as optim_mod
    rng = np.random.RandomState(100)
    obs = [(rng.randn(300, 3) * 0.12 + np.array([0.0, 0.05, -0.1])).astype(np.float32) for _ in range(3)]
    record = {'grad': [], 'z_before': [], 'lr': []}
    real_adam = optim_mod.Adam

    class RecordingAdam(real_adam):
        def step(self, closure=None):
            p = self.param_groups[0]['params'][0]
            record['grad'].append(p.grad.detach().clone().numpy().reshape(-1))
            record['z_before'].append(p.detach().clone().numpy().reshape(-1))
            record['lr'].append(self.param_groups[0]['lr'])
            return super().step(closure)

    optim_mod.Adam = RecordingAdam
    traj = {}
    try:
        n_iter = 12
        dec = make_ensemble(0, anchors)
        dec.train()                                                    # fitting_pointclouds.py:268
        lambdas = {'surface': 2.0, 'reg_global': 0.25, 'reg_unobserved': 10, 'reg_loc': 0.05, 'symm_dist': 5.0}
        traj['lambdas_initial'] = np.array([lambdas[k] for k in sorted(lambdas)], np.float64)
        schedule = {'lr': {200: 2, 400: 2, 600: 2, 800: 2}, 'symm_dist': {200: 10, 500: 9999},
                    'reg_glob': {200: 3, 600: 10}, 'reg_loc': {500: 3, 600: 10}}
        np.random.seed(0)
        torch.manual_seed(0)                                           # fitting_pointclouds.py:230-231
        z, anc = inference_identity_space(dec, [torch.from_numpy(o) for o in obs], lambdas,
                                          n_steps=n_iter * 100, schedule_cfg=schedule, step_scale=0.01)
    finally:
        optim_mod.Adam = real_adam
    traj['z_final'] = z.detach().numpy().reshape(-1)
    traj['anchors_final'] = anc.detach().numpy().reshape(39, 3)
    traj['lambdas_final'] = np.array([lambdas[k] for k in sorted(lambdas)], np.float64)
    traj['grads'] = np.stack(record['grad'])
    traj['z_before'] = np.stack(record['z_before'])
    traj['lrs'] = np.array(record['lr'], np.float64)
    print('fit', n_iter, 'iters: |z| =', float(z.detach().norm()), 'grad norms', np.linalg.norm(traj['grads'], axis=1)[:4])
    np.savez_compressed(os.path.join(HERE, 'fit_identity.npz'), obs=np.stack(obs), **traj)

    # ---------------------------------------------------------------- joint fitting (identity + expression, Broyden)
    # inference_iterative_root_finding_joint hard-codes `.cuda()` on two index tensors (fitting.py:72,137); to run the
    # UNMODIFIED function on this CPU-only box Tensor.cuda is patched to the identity for the duration of the call.
    rng = np.random.RandomState(200)
    obs_j = [(rng.randn(200, 3) * 0.1 + np.array([0.0, 0.05, -0.1])).astype(np.float32) for _ in range(3)]
    rec = {'grads': [], 'params': []}

    class RecordingAdam2(real_adam):
        def step(self, closure=None):
            p_ = self.param_groups[0]['params'][0]
            rec['grads'].append(p_.grad.detach().clone().numpy().copy())
            rec['params'].append(p_.detach().clone().numpy().copy())
            return super().step(closure)

    optim_mod.Adam = RecordingAdam2
    real_cuda = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self
    try:
        dec = make_ensemble(0, anchors)
        dec.train()
        torch.manual_seed(10)
        dfn = DeformationNetwork(mode='compress', lat_dim_expr=200, lat_dim_id=32, lat_dim_glob_shape=64,
                                 lat_dim_loc_shape=32, n_loc=39, anchors=anchors, hidden_dim=512, nlayers=6,
                                 out_dim=3, input_dim=3)
        dfn.eval()                                                      # fitting_pointclouds.py:229
        lambdas = {'surface': 2.0, 'reg_expr': 0.01, 'reg_global': 0.25, 'reg_unobserved': 10, 'reg_loc': 0.05,
                   'symm_dist': 5.0}                                    # fitting_pointclouds.py:253-259
        schedule = {'lr': {200: 2, 400: 2, 600: 2, 800: 2}, 'symm_dist': {200: 10, 500: 9999},
                    'reg_glob': {200: 3, 600: 10}, 'reg_loc': {500: 3, 600: 10}, 'reg_expr': {600: 10}}
        np.random.seed(0)
        torch.manual_seed(0)
        import io, contextlib
        with contextlib.redirect_stdout(io.StringIO()):
            z_ex, z_id, anc = inference_iterative_root_finding_joint(dec, dfn, [torch.from_numpy(o) for o in obs_j],
                                                                     lambdas, n_steps=400, schedule_cfg=schedule,
                                                                     step_scale=0.01)
    finally:
        optim_mod.Adam = real_adam
        torch.Tensor.cuda = real_cuda
    # opt.step() (identity) is called before opt_expr.step(): records alternate id, expr, id, expr ...
    np.savez_compressed(os.path.join(HERE, 'fit_joint.npz'), obs=np.stack(obs_j),
                        z_id_final=z_id.detach().numpy().reshape(-1), z_ex_final=z_ex.detach().numpy().reshape(3, 200),
                        anchors_final=anc.detach().numpy().reshape(39, 3),
                        grads_id=np.stack([g.reshape(-1) for g in rec['grads'][0::2]]),
                        grads_ex=np.stack([g.reshape(3, 200) for g in rec['grads'][1::2]]),
                        z_id_before=np.stack([g.reshape(-1) for g in rec['params'][0::2]]),
                        z_ex_before=np.stack([g.reshape(3, 200) for g in rec['params'][1::2]]))
    print('joint fit: 4 iterations, |z_id| %.4f |z_ex| %.4f, grad norms id %s' % (
        float(z_id.detach().norm()), float(z_ex.detach().norm()),
        [round(float(np.linalg.norm(g)), 4) for g in rec['grads'][0::2]]))


if __name__ == '__main__':
    main()

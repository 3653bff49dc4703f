#!/usr/bin/env python
"""Golden fixture for the Broyden correspondence search, from the UNMODIFIED reference
(src/NPHM/models/iterative_root_finding.py:91-168 `search`, multi_corresp=False as the joint fitter calls it).

    python tests/golden/make_golden_search.py          # build container only (needs /root/reference)

Two observations x 300 points against the seeded deformation network of make_golden.py (seed 10); its output layer is
scaled by 80 so that the search needs several quasi-Newton steps and ten samples fail to converge (all exit paths
of the loop are pinned), plus the multi-start variant (multi_corresp=True, seed 33) on the first 60 points of each
observation.  Writes search.npz: inputs, the reference's correspondences, residual norms, valid masks."""
import os
import numpy as np
import torch

import make_golden as G            # stubs the unused third-party imports and puts the reference on sys.path
from NPHM.models.deepSDF import DeformationNetwork
from NPHM.models.iterative_root_finding import search

OUT_SCALE = 80.0


def main():
    anchors64 = np.load(os.path.join(G.REF, 'assets', 'anchors_39.npy'))
    anchors = torch.from_numpy(anchors64).float().unsqueeze(0).unsqueeze(0)
    lat_mean = np.load(os.path.join(G.REF, 'assets', 'nphm_lat_mean.npy'))
    lat_std = np.load(os.path.join(G.REF, 'assets', 'nphm_lat_std.npy'))
    dec = G.make_ensemble(0, anchors)
    lat = G.sample_latent(1, torch.from_numpy(lat_mean), torch.from_numpy(lat_std))
    with torch.no_grad():
        _, anc = dec(torch.zeros(1, 1, 3), lat.reshape(1, 1, -1), None)
    torch.manual_seed(10)
    dfn = DeformationNetwork(mode='compress', lat_dim_expr=200, lat_dim_id=32, lat_dim_glob_shape=64,
                             lat_dim_loc_shape=32, n_loc=39, anchors=anchors, hidden_dim=512, nlayers=6,
                             out_dim=3, input_dim=3)
    dfn.eval()
    with torch.no_grad():
        dfn.defDeepSDF.lin6.weight.mul_(OUT_SCALE)
        dfn.defDeepSDF.lin6.bias.mul_(OUT_SCALE)
    torch.manual_seed(21)
    z_ex = torch.randn(2, 200) * 0.1
    rng = np.random.RandomState(22)
    obs = (rng.randn(2, 300, 3) * 0.15 + np.array([0.0, 0.05, -0.1])).astype(np.float32)
    n_point = obs.shape[1]
    glob_cond = torch.cat([lat.reshape(1, 1, -1).repeat(2, 1, 1), z_ex.unsqueeze(1)], dim=-1)
    sa = anc.clone().unsqueeze(1).repeat(2, n_point, 1, 1)
    xc, res = search(torch.from_numpy(obs), glob_cond.repeat(1, n_point, 1), dfn, sa, multi_corresp=False)
    valid = res['valid_ids'].numpy()
    diff = res['diff'].numpy()
    print('valid %d / %d, diff quantiles %s, max |xc-obs| %.4f' % (valid.sum(), valid.size,
          np.quantile(diff, [0.1, 0.5, 0.9, 1.0]), float((xc - torch.from_numpy(obs)).abs().max())))
    # multi-start variant (the function's default): 5 starts per point, perturbations drawn from torch's global generator
    n_multi = 60
    torch.manual_seed(33)
    # (the reference's multi-start residual only supports a batch of one: iterative_root_finding.py:143-146)
    xc_m, res_m = search(torch.from_numpy(obs[:1, :n_multi]), glob_cond[:1].repeat(1, n_multi, 1), dfn, sa[:1, :n_multi],
                         multi_corresp=True)
    print('multi-start: valid %d / %d' % (int(res_m['valid_ids'].sum()), res_m['valid_ids'].numel()))
    np.savez_compressed(os.path.join(G.HERE, 'search.npz'), out_scale=np.float32(OUT_SCALE), latent_id=lat.numpy(),
                        z_ex=z_ex.numpy(), anchors=anc.numpy().reshape(39, 3), obs=obs, xc=xc.detach().numpy(),
                        diff=diff.reshape(2, n_point), valid=valid.reshape(2, n_point),
                        multi_seed=np.int64(33), multi_n=np.int64(n_multi), xc_multi=xc_m.detach().numpy(),
                        valid_multi=res_m['valid_ids'].numpy(), diff_multi=res_m['diff'].numpy())


if __name__ == '__main__':
    main()

#!/usr/bin/env python
"""Golden fixture for the surface term of the joint fitter and its gradients, from the UNMODIFIED reference modules
(decoder in training mode as scripts/fitting/fitting_pointclouds.py:268 sets it; loss as in fitting.py:114-125):

    sdf = decoder(xc, z.repeat(nb, 1, 1), None)[0];  l = sdf[valid].abs();  loss = l[l < clamp].mean()

    python tests/golden/make_golden_surface.py          # build container only (needs /root/reference)

Writes surface_grad.npz: points, validity mask, latent, clamp, loss, d loss / d points, d loss / d latent (reference autograd)."""
import os
import numpy as np
import torch

import make_golden as G            # stubs the unused third-party imports and puts the reference on sys.path


def main():
    anchors64 = np.load(os.path.join(G.REF, 'assets', 'anchors_39.npy'))
    anchors = torch.from_numpy(anchors64).float().unsqueeze(0).unsqueeze(0)
    lat_mean = np.load(os.path.join(G.REF, 'assets', 'nphm_lat_mean.npy'))
    lat_std = np.load(os.path.join(G.REF, 'assets', 'nphm_lat_std.npy'))
    dec = G.make_ensemble(0, anchors)
    dec.train()
    z = G.sample_latent(3, torch.from_numpy(lat_mean), torch.from_numpy(lat_std)).reshape(1, 1, -1)
    torch.manual_seed(41)
    xc = torch.randn(2, 300, 3) * 0.15 + torch.tensor([0.0, 0.05, -0.1])
    valid = torch.rand(2, 300) > 0.2
    out = {}
    for clamp in (0.1, 0.02):
        xa = xc.clone().requires_grad_(True)
        za = z.clone().requires_grad_(True)
        sdf, _ = dec(xa, za.repeat(2, 1, 1), None)
        l = sdf[valid, :].abs()
        loss = l[l < clamp].mean()
        loss.backward()
        tag = '%g' % clamp
        out['loss_' + tag] = np.float32(loss.item())
        out['kept_' + tag] = np.int64(int((l < clamp).sum()))
        out['grad_points_' + tag] = xa.grad.numpy().copy()
        out['grad_latent_' + tag] = za.grad.reshape(-1).numpy().copy()
        print('clamp %s: loss %.6f kept %d |g_x| %.4e |g_z| %.4e' % (tag, loss.item(), int((l < clamp).sum()),
              float(xa.grad.abs().max()), float(za.grad.abs().max())))
    np.savez_compressed(os.path.join(G.HERE, 'surface_grad.npz'), points=xc.numpy(), valid=valid.numpy(),
                        latent=z.reshape(-1).numpy(), clamps=np.array([0.1, 0.02], np.float32), **out)


if __name__ == '__main__':
    main()

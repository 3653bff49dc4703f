#!/usr/bin/env python
"""Golden for the NPM baseline DeepSDF (scripts/configs/npm.yaml:2-4: lat_dim 512, hidden 1024, 8 layers) from the
UNMODIFIED reference (build container only):   python tests/golden/make_golden_npm.py
Weights are not stored: the drop-in class reproduces the reference's initialisation from the same seed (sha256 below)."""
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
    sys.modules[name] = types.ModuleType(name)
from NPHM.models.deepSDF import DeepSDF          # noqa: E402


def main():
    torch.manual_seed(12)
    net = DeepSDF(lat_dim=512, hidden_dim=1024, nlayers=8, geometric_init=True)
    h = hashlib.sha256()
    sd = net.state_dict()
    for k in sorted(sd):
        h.update(k.encode()); h.update(np.ascontiguousarray(sd[k].numpy()).tobytes())
    rng = np.random.RandomState(5)
    pts = ((rng.rand(2, 700, 3) - 0.5) * 1.2).astype(np.float32)
    lat = torch.from_numpy(np.load(os.path.join(REF, 'assets', 'npm_lat_mean.npy')).astype(np.float32))
    std = torch.from_numpy(np.load(os.path.join(REF, 'assets', 'npm_lat_std.npy')).astype(np.float32))
    torch.manual_seed(13)
    codes = torch.stack([lat + 0.85 * std * torch.randn(512), lat - 0.5 * std * torch.randn(512)])      # 2 x 512
    with torch.no_grad():
        out, _ = net(torch.from_numpy(pts), codes[:, None, :].repeat(1, pts.shape[1], 1))
    # second net: default (non-geometric) initialisation, 3 outputs - a less degenerate function of the inputs
    torch.manual_seed(14)
    net2 = DeepSDF(lat_dim=512, hidden_dim=1024, nlayers=8, geometric_init=False, out_dim=3)
    with torch.no_grad():
        out2, _ = net2(torch.from_numpy(pts), (codes * 5.0)[:, None, :].repeat(1, pts.shape[1], 1))
    print('plain net: range', float(out2.min()), float(out2.max()))
    np.savez_compressed(os.path.join(HERE, 'npm.npz'), points=pts, codes=codes.numpy(), sdf=out.numpy(), out_plain=out2.numpy(),
                        sha256=np.array(h.hexdigest()))
    print('npm golden: sdf range', float(out.min()), float(out.max()), 'codes norm', float(codes.norm(dim=1).mean()))


if __name__ == '__main__':
    main()

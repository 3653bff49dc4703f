"""ORACLE - test infrastructure only: imports the UNMODIFIED reference modules of the hot path.

Source of the modules, in this order: ``oracle/_ref/`` (byte-identical copies made by ``oracle/make_ref.py`` in the
build container; they travel to the GPU box), then ``/root/reference`` (build container only).  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s reference / cpu_baseline / stock_gpu legs may import this module.

The reference imports four third-party packages at module top level that the path never uses and that are not in
this image (``trimesh``, ``mcubes``, ``pyvista``, ``pytorch3d.ops``): they are stubbed for the duration of the import.
The reference package is imported under its own name ``NPHM`` (its modules import each other absolutely,
``fitting.py:10-11``) and then REMOVED from ``sys.modules`` again, so that it never collides with
``nphm_b200.install_as_nphm()`` in the same process; the returned namespace keeps the module objects alive.
"""
from __future__ import annotations

import importlib
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_CANDIDATES = (os.path.join(HERE, '_ref'), '/root/reference')
_STUBS = ('trimesh', 'mcubes', 'pyvista', 'pytorch3d', 'pytorch3d.ops')
_MODULES = ('models.EnsembledDeepSDF', 'models.deepSDF', 'models.diff_operators', 'models.iterative_root_finding',
            'models.reconstruction', 'models.fitting', 'utils.reconstruction')
_cache = None


class ReferenceUnavailable(RuntimeError):
    pass


def reference_root() -> str:
    for root in _CANDIDATES:
        if os.path.exists(os.path.join(root, 'src', 'NPHM', 'models', 'EnsembledDeepSDF.py')):
            return root
    raise ReferenceUnavailable('no reference checkout: run `python oracle/make_ref.py` where /root/reference exists')


def available() -> bool:
    try:
        reference_root()
        return True
    except ReferenceUnavailable:
        return False


def load():
    """-> namespace with the reference's modules (``EnsembledDeepSDF``, ``deepSDF``, ``reconstruction``, ``fitting``,
    ``iterative_root_finding``, ``diff_operators``, ``utils_reconstruction``), ``root`` and ``assets`` (dict of arrays)."""
    global _cache
    if _cache is not None:
        return _cache
    root = reference_root()
    saved = {k: v for k, v in sys.modules.items() if k == 'NPHM' or k.startswith('NPHM.') or k in _STUBS}
    for k in saved:
        del sys.modules[k]
    for name in _STUBS:
        sys.modules[name] = types.ModuleType(name)
    sys.modules['pytorch3d.ops'].knn_points = None
    sys.modules['pytorch3d.ops'].knn_gather = None
    sys.modules['pytorch3d'].ops = sys.modules['pytorch3d.ops']
    src = os.path.join(root, 'src')
    sys.path.insert(0, src)
    ns = types.SimpleNamespace(root=root)
    try:
        for name in _MODULES:
            mod = importlib.import_module('NPHM.' + name)
            setattr(ns, name.split('.')[-1] if not name.startswith('utils.') else 'utils_reconstruction', mod)
        # the identity-space training losses: optional, older prebuilt copies do not carry the file
        ns.loss_functions = None
        if os.path.exists(os.path.join(src, 'NPHM', 'models', 'loss_functions.py')):
            ns.loss_functions = importlib.import_module('NPHM.models.loss_functions')
    finally:
        sys.path.remove(src)
        for k in [k for k in sys.modules if k == 'NPHM' or k.startswith('NPHM.') or k in _STUBS]:
            del sys.modules[k]
        sys.modules.update(saved)
    ns.assets = {n: np.load(os.path.join(root, 'assets', n + '.npy'))
                 for n in ('anchors_39', 'nphm_lat_mean', 'nphm_lat_std')}
    _cache = ns
    return ns


# ---------------------------------------------------------------------------------------- seeded models
def mean_anchors(ns):
    import torch
    return torch.from_numpy(ns.assets['anchors_39']).float().unsqueeze(0).unsqueeze(0)     # fitting_pointclouds.py:80


def make_ensemble(ns, seed=0, device='cpu'):
    """The reference's FastEnsembleDeepSDFMirrored with the NPHM config (scripts/configs/nphm.yaml:2-7), default init."""
    import torch
    torch.manual_seed(seed)
    dec = ns.EnsembledDeepSDF.FastEnsembleDeepSDFMirrored(
        lat_dim_glob=64, lat_dim_loc=32, n_loc=39, n_symm_pairs=16, anchors=mean_anchors(ns), hidden_dim=200,
        n_layers=4, pos_mlp_dim=256)
    dec = dec.to(device)
    dec.anchors = dec.anchors.to(device)          # plain attribute, EnsembledDeepSDF.py:192
    return dec


def make_deformation(ns, seed=10, device='cpu'):
    import torch
    torch.manual_seed(seed)
    dfn = ns.deepSDF.DeformationNetwork(mode='compress', lat_dim_expr=200, lat_dim_id=32, lat_dim_glob_shape=64,
                                        lat_dim_loc_shape=32, n_loc=39, anchors=mean_anchors(ns), hidden_dim=512,
                                        nlayers=6, out_dim=3, input_dim=3)
    dfn.eval()
    return dfn.to(device)


def sample_latent(ns, seed):
    import torch
    torch.manual_seed(seed)
    mean, std = torch.from_numpy(ns.assets['nphm_lat_mean']), torch.from_numpy(ns.assets['nphm_lat_std'])
    return torch.randn(mean.shape) * std * 0.85 + mean                                      # fitting_pointclouds.py:206

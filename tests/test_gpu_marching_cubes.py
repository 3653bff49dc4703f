"""GPU parity tests for the marching-cubes kernels: bit-exact (vertex order, fp64 positions, triangle ids) against
the oracle's sequential restatement, plus size-independent properties at the full 256^3 size."""
import numpy as np
import pytest
import torch

from conftest import MAXI, MINI, mesh_edge_stats, noise_volume, sphere_volume
from oracle import nphm_oracle as O

pytestmark = pytest.mark.gpu


def _gpu_mc(vol, iso=0.0, negate=False, device='cuda:0'):
    from nphm_b200 import _native
    v, t = _native.marching_cubes_device(torch.from_numpy(np.ascontiguousarray(vol)).to(device), iso, negate)
    return v.cpu().numpy(), t.cpu().numpy()


@pytest.mark.parametrize('shape,seed', [((17, 17, 17), 7), ((9, 33, 5), 8), ((2, 2, 2), 9), ((40, 3, 21), 10),
                                        ((5, 2, 64), 11), ((2, 70, 70), 12)])
def test_noise_volumes_bit_exact(cuda_device, shape, seed):
    """uniform(-1,1) noise exercises all 256 cases, boundary ownership rules and ragged dimensions."""
    vol = noise_volume(shape, seed)
    for iso, negate in ((0.0, False), (0.25, False), (0.0, True)):
        rv, rt = O.marching_cubes(vol, iso, negate)
        gv, gt = _gpu_mc(vol, iso, negate)
        assert gv.shape == rv.shape and gt.shape == rt.shape
        assert np.array_equal(gt.astype(np.uint64), rt)
        assert np.array_equal(gv, rv)                   # fp64 positions bit for bit


def test_exact_zero_samples_and_equal_corners(cuda_device):
    """samples exactly at the iso value (<= classification) and f1 == f2 edges (midpoint rule)."""
    vol = noise_volume((12, 12, 12), 3)
    vol[vol > 0.5] = 0.0
    vol[3:5, 3:5, 3:5] = 0.0
    vol = np.round(vol * 4) / 4
    rv, rt = O.marching_cubes(vol, 0.0)
    gv, gt = _gpu_mc(vol, 0.0)
    assert np.array_equal(gt.astype(np.uint64), rt) and np.array_equal(gv, rv)


def test_degenerate_and_empty(cuda_device):
    from nphm_b200 import _native
    for shape in ((1, 8, 8), (8, 1, 8), (8, 8, 1)):
        v, t = _gpu_mc(noise_volume(shape, 1))
        assert v.shape == (0, 3) and t.shape == (0, 3)
    v, t = _gpu_mc(np.ones((6, 6, 6), np.float32))          # no crossing at all
    assert len(v) == 0 and len(t) == 0
    v, t = _native.marching_cubes_host(sphere_volume(10), 0.0)      # host-buffer entry point
    rv, rt = O.marching_cubes(sphere_volume(10), 0.0)
    assert np.array_equal(v, rv) and np.array_equal(t, rt)


def test_sharded_slabs_reassemble_to_the_whole(cuda_device):
    """x-slabs with a one-layer ghost reproduce the single-volume numbering exactly (config 5 plumbing)."""
    from nphm_b200 import _native
    vol = noise_volume((21, 10, 13), 5)
    rv, rt = O.marching_cubes(vol, 0.0)
    dev_vol = torch.from_numpy(vol).to(cuda_device)
    bounds = [0, 6, 7, 15, 20]             # cell-layer ranges per shard (20 cell layers)
    counts, keep = [], []
    for lo, hi in zip(bounds[:-1], bounds[1:]):
        ghost = lo > 0
        slab = dev_vol[lo - (1 if ghost else 0): hi + 1].contiguous()
        nv, nt, p, ws = _native.marching_cubes_count(slab, 0.0, False, x_global0=lo - (1 if ghost else 0), ghost_lo=ghost)
        counts.append((nv, nt))
        keep.append((slab, p, ws))
    base = 0
    verts, tris = [], []
    for (nv, nt), (slab, p, ws) in zip(counts, keep):
        v, t = _native.marching_cubes_emit(slab, p, ws, nv, nt, base)
        verts.append(v.cpu().numpy()); tris.append(t.cpu().numpy())
        base += nv
    assert np.array_equal(np.concatenate(verts), rv)
    assert np.array_equal(np.concatenate(tris).astype(np.uint64), rt)


def test_mesh_from_logits_dropin(cuda_device):
    from nphm_b200.utils.reconstruction import mesh_from_logits
    vol = sphere_volume(32)
    a, b = vol.reshape(-1).copy(), vol.reshape(-1).copy()
    mesh = mesh_from_logits(a, MINI, MAXI, 32)
    rv, rt = O.mesh_from_logits(b, MINI, MAXI, 32)
    assert np.array_equal(a, b)                                       # both negated in place
    assert np.array_equal(np.asarray(mesh.vertices), rv)
    assert np.array_equal(np.asarray(mesh.faces).astype(np.uint64), rt)


def test_full_size_256_properties_and_parity(cuda_device):
    """256^3 (BASELINE config 2 size): bit-exact against the oracle and closed-manifold / Euler properties."""
    vol = sphere_volume(256, radius=0.37)
    gv, gt = _gpu_mc(vol, 0.0)
    dup, unmatched = mesh_edge_stats(gt)
    assert dup == 0 and unmatched == 0
    assert len(gv) - 3 * len(gt) // 2 + len(gt) == 2
    world = gv / 255.0 - 0.5
    r = np.linalg.norm(world - np.array([0.03, -0.02, 0.01]), axis=1)
    assert np.abs(r - 0.37).max() < 1e-4
    rv, rt = O.marching_cubes(vol, 0.0)
    assert np.array_equal(gt.astype(np.uint64), rt) and np.array_equal(gv, rv)

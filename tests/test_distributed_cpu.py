"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: slab planning, count exchange, id offsets, gather.
The marching-cubes kernels are replaced by the oracle's sequential restatement applied per slab (ghost layer
emulated by running the oracle on the planes up to the slab and keeping only the new part)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import noise_volume


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


def test_plan_slabs_covers_all_layers():
    from nphm_b200.distributed import plan_slabs, slab_planes
    for res, world in ((256, 8), (512, 8), (20, 3), (5, 8), (2, 2)):
        plan = plan_slabs(res, world)
        assert plan[0][0] == 0 and plan[-1][1] == res - 1
        assert all(a[1] == b[0] for a, b in zip(plan[:-1], plan[1:]))
        sizes = [c1 - c0 for c0, c1 in plan]
        assert max(sizes) - min(sizes) <= 1
        for c0, c1 in plan:
            p0, n, ghost = slab_planes(c0, c1)
            if c1 > c0:
                assert p0 == max(c0 - 1, 0) and p0 + n == c1 + 1 and ghost == (c0 > 0)
            else:
                assert n == 0


def _oracle_slab_mc(vol_full):
    """CPU stand-ins for marching_cubes_count / emit with the same slab semantics as the CUDA kernels."""
    from oracle import nphm_oracle as O

    def count(vol, iso, negate, x_global0=0, ghost_lo=False):
        upto = x_global0 + vol.shape[0]
        v_all, t_all = O.marching_cubes(vol_full[:upto], iso, negate)
        first_own_layer = x_global0 + (1 if ghost_lo else 0)
        if first_own_layer > 0:
            v_prev, t_prev = O.marching_cubes(vol_full[:first_own_layer + 1], iso, negate)
        else:
            v_prev, t_prev = np.zeros((0, 3)), np.zeros((0, 3), np.uint64)
        # sequential numbering: everything the earlier layers created comes first
        assert np.array_equal(v_all[:len(v_prev)], v_prev) and np.array_equal(t_all[:len(t_prev)], t_prev)
        own_v, own_t = v_all[len(v_prev):], t_all[len(t_prev):]
        return len(own_v), len(own_t), (own_v, own_t, len(v_prev)), None

    def emit(vol, params, ws, nv, nt, base):
        own_v, own_t, n_prev = params
        assert base == n_prev                               # id base from the count exchange == sequential numbering
        return torch.from_numpy(own_v.copy()), torch.from_numpy(own_t.astype(np.int64))

    return count, emit


def _worker(rank, world, port, res, q):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        from nphm_b200.distributed import exchange_counts, extract_mesh_sharded
        from oracle import nphm_oracle as O
        vol_full = noise_volume((res, 6, 7), seed=21)
        counts = exchange_counts(10 + rank, 20 + rank)
        assert counts == [(10 + r, 20 + r) for r in range(world)]
        count, emit = _oracle_slab_mc(vol_full)
        verts, tris = extract_mesh_sharded(lambda p0, n: torch.from_numpy(vol_full[p0:p0 + n]), res, 0.0, False,
                                           mc_count=count, mc_emit=emit)
        if rank == 0:
            rv, rt = O.marching_cubes(vol_full, 0.0)
            ok = np.array_equal(verts.numpy(), rv) and np.array_equal(tris.numpy().astype(np.uint64), rt)
            q.put(bool(ok))
        else:
            assert verts is None and tris is None
    finally:
        dist.destroy_process_group()


def test_sharded_extraction_two_ranks_gloo():
    ctx = mp.get_context('spawn')
    for world, res in ((2, 9), (2, 2)):
        q = ctx.SimpleQueue()
        port = _free_port()
        procs = [ctx.Process(target=_worker, args=(r, world, port, res, q)) for r in range(world)]
        for p in procs:
            p.start()
        for p in procs:
            p.join(180)
            assert p.exitcode == 0
        assert q.get() is True


# ---------------------------------------------------------------------------------------------- point-sharded fitting
class _TorchShardOps:
    """CPU stand-in for the native per-rank kernels of inference_identity_space_sharded: autograd through the composite
    decoder for the surface term, autograd for the regularisers, the oracle's Adam."""

    def __init__(self, dec):
        self.dec = dec
        self.latent = torch.zeros(dec.lat_dim)
        self.m = np.zeros(dec.lat_dim, np.float32)
        self.v = np.zeros(dec.lat_dim, np.float32)
        self.first_grad = None

    def surface(self, pts, clamp):
        if pts.shape[0] == 0:
            return torch.zeros_like(self.latent), torch.tensor(float('nan')), torch.tensor(0.0)
        z = self.latent.clone().requires_grad_(True)
        sdf, _ = self.dec(pts[None], z.reshape(1, 1, -1), None)
        l = sdf.abs()
        keep = l < clamp
        if int(keep.sum()) == 0:
            return torch.zeros_like(self.latent), torch.tensor(float('nan')), torch.tensor(0.0)
        loss = l[keep].mean()
        (g,) = torch.autograd.grad(loss, z)
        return g, loss.detach(), keep.sum().float()

    def apply(self, mean_grad, stats, lambdas, clamp, lr, step):
        from nphm_b200.models.fitting import _latent_regularisers
        from oracle import nphm_oracle as O
        if self.first_grad is None:
            self.first_grad = (mean_grad.numpy().copy(), stats.numpy().copy())
        z = self.latent.clone().requires_grad_(True)
        regs = _latent_regularisers(self.dec, z.reshape(1, 1, -1))
        reg_loss = sum(regs[k] * lambdas[k] for k in lambdas if k != 'surface')
        (g_reg,) = torch.autograd.grad(reg_loss, z)
        g = (lambdas['surface'] * mean_grad + g_reg).numpy().astype(np.float32)
        zn, self.m, self.v = O.adam_step(self.latent.numpy().copy(), g, self.m, self.v, step, float(lr))
        self.latent = torch.from_numpy(np.asarray(zn, np.float32))


def _fit_worker(rank, world, port, q):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        import nphm_b200.models.fitting as F
        from nphm_b200.distributed import inference_identity_space_sharded, shard_rows
        from conftest import make_ensemble
        from fit_common import golden_fit_setup
        F.NUM_POINTS_PER_OBSERVATION = 37                 # 5 x 37 = 185 points per iteration: ragged over 2 ranks
        assert [shard_rows(185, 2, r) for r in range(2)] == [(0, 93), (93, 185)]
        _, obs, lambdas, schedule = golden_fit_setup()
        dec = make_ensemble(0).train()
        np.random.seed(0)
        torch.manual_seed(0)
        torch.set_num_threads(2)
        ops = _TorchShardOps(dec)
        z, anchors = inference_identity_space_sharded(dec, obs, lambdas, n_steps=400, schedule_cfg=schedule, step_scale=0.01,
                                                      ops=ops)
        assert z.shape == (1, 1, 1344) and anchors.shape == (1, 39, 3)
        q.put((rank, (z.detach().numpy().reshape(-1).copy(),) + ops.first_grad))
    finally:
        dist.destroy_process_group()


def test_point_sharded_fitting_two_ranks_gloo():
    """4 iterations of the sharded identity fit on 2 gloo ranks: both ranks end with the SAME latent (replicated update, no
    broadcast) and it equals the 1-rank run of the same loop up to summation order."""
    ctx = mp.get_context('spawn')
    results = {}
    for world in (2, 1):
        q = ctx.SimpleQueue()
        port = _free_port()
        procs = [ctx.Process(target=_fit_worker, args=(r, world, port, q)) for r in range(world)]
        for p in procs:
            p.start()
        out = [q.get() for _ in range(world)]
        for p in procs:
            p.join(300)
            assert p.exitcode == 0
        results[world] = dict(out)
    z2a, g2a, st2a = results[2][0]
    z2b, g2b, st2b = results[2][1]
    z1, g1, st1 = results[1][0]
    assert np.array_equal(z2a, z2b) and np.array_equal(g2a, g2b)          # replicated state stays bit-identical
    assert np.abs(z2a).max() > 1e-3                                       # the latent moved
    # the one all-reduce reproduces the unsharded mean gradient / kept count / loss sum of the first iteration
    assert st2a[0] == st1[0] and abs(st2a[1] - st1[1]) < 1e-5 * abs(st1[1])
    assert np.abs(g2a - g1).max() < 1e-5 * np.abs(g1).max()
    # later iterations: == the unsharded loop up to fp32 summation order; Adam's first steps are sign-like, so elements
    # whose gradient is at round-off level land one step apart (cf. test_autograd_fallback_follows_reference)
    close = np.abs(z2a - z1) < 2e-5
    assert close.mean() > 0.9, (close.mean(), np.abs(z2a - z1).max())

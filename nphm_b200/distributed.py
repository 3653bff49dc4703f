"""Multi-GPU mesh extraction: the grid is partitioned in contiguous x-slabs, one per rank.

  * SDF query: every point is independent -> no collective.  A rank evaluates the planes its cells touch (its own
    cell layers, the closing plane, and one ghost cell layer below so that vertices shared with the previous rank
    resolve to that rank's ids); the <= 2 extra planes are recomputed rather than exchanged.
  * Marching cubes: one exchange step - ``all_gather`` of the per-rank (vertex, triangle) counts gives every rank the
    id base of its first vertex, so the concatenated result is IDENTICAL (ids, order, fp64 positions) to a single-GPU run.
  * The triangle / vertex buffers are collected on rank 0 in one batched point-to-point exchange (variable sizes; every
    buffer lands directly in its slice of the result).

One process per GPU, ``torch.distributed`` (NCCL on GPUs; the host logic is exercised with gloo on CPU in
tests/test_distributed_cpu.py).
"""
from __future__ import annotations

from typing import Callable, List, Optional, Tuple

import torch
import torch.distributed as dist


def plan_slabs(res: int, world: int) -> List[Tuple[int, int]]:
    """Cell layers [c0, c1) of each rank; res-1 cell layers split as evenly as possible (first ranks get the extra)."""
    n_layers = res - 1
    base, extra = divmod(n_layers, world)
    out, c = [], 0
    for r in range(world):
        n = base + (1 if r < extra else 0)
        out.append((c, c + n))
        c += n
    return out


def slab_planes(c0: int, c1: int) -> Tuple[int, int, bool]:
    """(first plane, number of planes, has ghost layer) a rank must evaluate for cell layers [c0, c1)."""
    if c1 <= c0:
        return c0, 0, False
    ghost = c0 > 0
    p0 = c0 - (1 if ghost else 0)
    return p0, c1 + 1 - p0, ghost


def exchange_counts(n_verts: int, n_tris: int, group=None, device='cpu') -> List[Tuple[int, int]]:
    world = dist.get_world_size(group)
    mine = torch.tensor([n_verts, n_tris], dtype=torch.int64, device=device)
    allc = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(allc, mine, group=group)
    return [(int(c[0]), int(c[1])) for c in allc]


def gather_mesh(verts: torch.Tensor, tris: torch.Tensor, counts: List[Tuple[int, int]], group=None, dst: int = 0):
    """Concatenate per-rank (verts (v,3) f64, tris (t,3) i64) on ``dst`` in rank order: ONE batched exchange
    (``batch_isend_irecv`` = one NCCL group): every sender posts its two buffers, ``dst`` receives each of them
    directly into its slice of the result at the prefix offset the count exchange gave (no staging buffers, no
    serial receive loop).  Returns (verts, tris) on ``dst`` and (None, None) elsewhere."""
    rank = dist.get_rank(group)
    world = dist.get_world_size(group)
    ops = []
    if rank != dst:
        if verts.shape[0]:
            ops.append(dist.P2POp(dist.isend, verts.contiguous(), dst, group))
        if tris.shape[0]:
            ops.append(dist.P2POp(dist.isend, tris.contiguous(), dst, group))
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()
        return None, None
    tot_v = sum(c[0] for c in counts)
    tot_t = sum(c[1] for c in counts)
    all_v = torch.empty(tot_v, 3, dtype=verts.dtype, device=verts.device)
    all_t = torch.empty(tot_t, 3, dtype=tris.dtype, device=tris.device)
    ov = ot = 0
    for r in range(world):
        nv, nt = counts[r]
        if r == dst:
            all_v[ov:ov + nv] = verts
            all_t[ot:ot + nt] = tris
        else:
            if nv:
                ops.append(dist.P2POp(dist.irecv, all_v[ov:ov + nv], r, group))      # contiguous row slice
            if nt:
                ops.append(dist.P2POp(dist.irecv, all_t[ot:ot + nt], r, group))
        ov += nv
        ot += nt
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()
    return all_v, all_t


class _Lap:
    """Device-side lap timer (CUDA events on the current stream); a no-op on CPU tensors / when disabled."""

    def __init__(self, sink: Optional[dict], device):
        self.sink = sink
        self.on = sink is not None and torch.device(device).type == 'cuda'
        self.marks = []
        if self.on:
            self.mark(None)

    def mark(self, name):
        if not self.on:
            return
        ev = torch.cuda.Event(enable_timing=True)
        ev.record()
        self.marks.append((name, ev))

    def close(self):
        if not self.on:
            return
        torch.cuda.synchronize()
        for (_, a), (name, b) in zip(self.marks[:-1], self.marks[1:]):
            self.sink[name] = self.sink.get(name, 0.0) + a.elapsed_time(b)


def extract_mesh_sharded(slab_volume_fn: Callable[[int, int], torch.Tensor], res: int, iso: float = 0.0,
                         negate: bool = True, group=None,
                         mc_count: Optional[Callable] = None, mc_emit: Optional[Callable] = None,
                         timings: Optional[dict] = None):
    """Sharded ``get_logits`` + marching cubes.

    ``slab_volume_fn(first_plane, n_planes)`` returns this rank's SDF planes as a (n_planes, res, res) float32 tensor.
    ``mc_count`` / ``mc_emit`` default to the CUDA kernels (``nphm_b200._native``); tests inject CPU stand-ins.
    ``timings`` (optional dict) accumulates device milliseconds per phase: ``sdf_ms``, ``mc_count_ms``,
    ``count_exchange_ms``, ``mc_emit_ms``, ``gather_ms``, and ``gather_bytes`` (what this rank sent over NVLink).
    Returns (verts, tris) in global index units on rank 0, (None, None) elsewhere."""
    if mc_count is None or mc_emit is None:
        from . import _native
        mc_count, mc_emit = _native.marching_cubes_count, _native.marching_cubes_emit
    rank = dist.get_rank(group)
    world = dist.get_world_size(group)
    c0, c1 = plan_slabs(res, world)[rank]
    p0, n_planes, ghost = slab_planes(c0, c1)
    cuda = dist.get_backend(group) == 'nccl'
    device = torch.device('cuda', torch.cuda.current_device()) if cuda else torch.device('cpu')
    lap = _Lap(timings, device)
    if n_planes:
        vol = slab_volume_fn(p0, n_planes)
        lap.mark('sdf_ms')
        nv, nt, params, ws = mc_count(vol, iso, negate, x_global0=p0, ghost_lo=ghost)
        device = vol.device
    else:
        lap.mark('sdf_ms')
        vol, nv, nt, params, ws = None, 0, 0, None, None
    lap.mark('mc_count_ms')
    counts = exchange_counts(nv, nt, group, device=device)
    lap.mark('count_exchange_ms')
    base = sum(c[0] for c in counts[:rank])
    if n_planes:
        verts, tris = mc_emit(vol, params, ws, nv, nt, base)
    else:
        verts = torch.empty(0, 3, dtype=torch.float64, device=device)
        tris = torch.empty(0, 3, dtype=torch.int64, device=device)
    lap.mark('mc_emit_ms')
    out = gather_mesh(verts, tris, counts, group)
    lap.mark('gather_ms')
    lap.close()
    if timings is not None:
        timings['gather_bytes'] = 0 if rank == 0 else int(verts.numel() * 8 + tris.numel() * 8)
    return out


def ensemble_slab_fn(decoder, latent: torch.Tensor, mini, maxi, res: int, nbatch_points: int):
    """``slab_volume_fn`` for the identity ensemble: in-kernel grid generation over this rank's planes; the eval quirk
    follows the GLOBAL flat index so the result does not depend on the number of ranks."""
    eng = decoder.engine()

    def fn(first_plane: int, n_planes: int) -> torch.Tensor:
        period = 0 if decoder.training else int(nbatch_points)
        out, _ = eng.query_grid(latent, mini, maxi, res, first_plane * res * res, n_planes * res * res, period)
        return out.view(n_planes, res, res)

    return fn

#!/usr/bin/env python
"""BASELINE.json configs[4]: one head, res^3 grid sharded in x-slabs over N GPUs, NCCL gather of the mesh on rank 0.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        tools/bench_sharded.py --res 512 [--check]

Strong scaling of a single mesh extraction; prints one JSON line on rank 0.  --check recomputes the mesh on rank 0
alone and verifies that the sharded result is identical (ids, order, fp64 positions).
"""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--res', type=int, default=512)
    ap.add_argument('--steps', type=int, default=3)
    ap.add_argument('--warmup', type=int, default=1)
    ap.add_argument('--check', action='store_true')
    args = ap.parse_args()
    from conftest import MAXI, MINI, make_ensemble, sample_latent
    from nphm_b200 import _native
    from nphm_b200.distributed import ensemble_slab_fn, extract_mesh_sharded

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    dist.init_process_group('nccl', device_id=dev)
    dec = make_ensemble(0, device=dev).eval()
    lat = sample_latent(1).to(dev)                     # the same head on every rank
    fn = ensemble_slab_fn(dec, lat, MINI, MAXI, args.res, 25000)

    def run():
        return extract_mesh_sharded(fn, args.res, 0.0, True)

    for _ in range(args.warmup):
        run()
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        verts, tris = run()
    e1.record()
    dist.barrier(); torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / args.steps], device=dev)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ok = None
    if args.check and rank == 0:
        vol, _ = dec.engine().query_grid(lat, MINI, MAXI, args.res, 0, args.res ** 3, 25000)
        v1, t1 = _native.marching_cubes_device(vol.view(args.res, args.res, args.res), 0.0, negate=True)
        ok = bool(torch.equal(v1, verts) and torch.equal(t1, tris))
    if rank == 0:
        print(json.dumps({'metric': 'sharded_mesh_extraction', 'res': args.res, 'n_gpus': world, 'ms_per_mesh': ms.item(),
                          'points_per_s': args.res ** 3 / (ms.item() * 1e-3), 'vertices': int(verts.shape[0]),
                          'triangles': int(tris.shape[0]), 'identical_to_single_gpu': ok, 'scaling': 'strong'}))
    dist.destroy_process_group()


if __name__ == '__main__':
    main()

#!/usr/bin/env python
"""BASELINE.json configs[3]: identity fitting of a batch of scans, ONE SCAN PER GPU (replicas only, no collective).

    python tools/bench_fit.py --steps 200                                   # 1 GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port P tools/bench_fit.py

Every rank fits its own synthetic scan (3 observations x 2500 points, seed 100 + rank) with the reference's loop
(`inference_identity_space`, 5 x 1000 sampled points per iteration, reference schedule) on the fused fitting kernels.
Prints one JSON line on rank 0: iterations/s and scans/hour over all ranks (time = max over ranks)."""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np
import torch
import torch.distributed as dist


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--steps', type=int, default=1000)
    ap.add_argument('--sharded', action='store_true',
                    help='ONE head, the sampled points of every iteration sharded over the ranks, one all-reduce per iteration '
                         '(nphm_b200.distributed.inference_identity_space_sharded) instead of one scan per GPU')
    args = ap.parse_args()
    from conftest import make_ensemble
    from nphm_b200.models.fitting import inference_identity_space
    world = int(os.environ.get('WORLD_SIZE', '1')); rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    dec = make_ensemble(0, device=dev).train()
    rng = np.random.RandomState(100 + (0 if args.sharded else rank))
    obs = [torch.from_numpy((rng.randn(2500, 3) * 0.12 + np.array([0.0, 0.05, -0.1])).astype(np.float32)).to(dev) for _ in range(3)]
    lambdas = {'surface': 2.0, 'reg_global': 0.25, 'reg_unobserved': 10, 'reg_loc': 0.05, 'symm_dist': 5.0}
    schedule = {'lr': {200: 2, 400: 2, 600: 2, 800: 2}, 'symm_dist': {200: 10, 500: 9999}, 'reg_glob': {200: 3, 600: 10},
                'reg_loc': {500: 3, 600: 10}}
    if args.sharded:
        from nphm_b200.distributed import inference_identity_space_sharded
        if world == 1 and not dist.is_initialized():
            if 'MASTER_ADDR' in os.environ and 'RANK' in os.environ:
                dist.init_process_group('nccl', device_id=dev)
            else:
                dist.init_process_group('nccl', device_id=dev, init_method='tcp://127.0.0.1:29533', rank=0, world_size=1)

        def fit(n):
            return inference_identity_space_sharded(dec, obs, dict(lambdas), n_steps=n, schedule_cfg=schedule)
    else:
        def fit(n):
            return inference_identity_space(dec, obs, dict(lambdas), n_steps=n, schedule_cfg=schedule)
    np.random.seed(0); torch.manual_seed(0)
    fit(5)                                                                                        # warm-up
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    np.random.seed(0); torch.manual_seed(0)
    z, _ = fit(args.steps)
    torch.cuda.synchronize(); dt = torch.tensor([time.perf_counter() - t0], device=dev)
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(json.dumps({'metric': 'identity_fit', 'n_gpus': world, 'iterations': args.steps, 's_per_scan': dt.item(),
                          'iters_per_s_total': world * args.steps / dt.item(), 'scans_per_hour': world * 3600 / dt.item(),
                          'finite': bool(torch.isfinite(z).all()),
                          'scaling': ('strong: one head, points sharded, 1 all-reduce of lat_dim + 2 floats per iteration; '
                                      'iters_per_s_total / n_gpus = iterations/s of that head') if args.sharded
                          else 'replicas (one scan per GPU, no collective)'}))
    if dist.is_initialized():
        dist.destroy_process_group()


if __name__ == '__main__':
    main()

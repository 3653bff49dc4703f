#!/usr/bin/env python
"""Joint identity + expression fitting (reference `inference_iterative_root_finding_joint`): iteration rate
and, with --profile, where one iteration's time goes (torch profiler, top CUDA ops + number of launches).

    python tools/bench_joint.py --steps 50 [--profile] [--obs 3]

Synthetic scan: `--obs` observations x 2500 points; every iteration samples 5 x 1000 points like the reference."""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np
import torch


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--steps', type=int, default=50)
    ap.add_argument('--obs', type=int, default=3)
    ap.add_argument('--profile', action='store_true')
    args = ap.parse_args()
    from conftest import make_ensemble, make_deformation
    from nphm_b200.models.fitting import inference_iterative_root_finding_joint
    dev = torch.device('cuda', 0)
    dec = make_ensemble(0, device=dev).train()               # fitting_pointclouds.py:268
    dfn = make_deformation(device=dev)
    with torch.no_grad():                                   # small deformations, like a trained field near the neutral pose
        dfn.defDeepSDF.lin6.weight.mul_(0.05); dfn.defDeepSDF.lin6.bias.mul_(0.05)
    rng = np.random.RandomState(7)
    obs = [torch.from_numpy((rng.randn(2500, 3) * 0.12 + np.array([0.0, 0.05, -0.1])).astype(np.float32)).to(dev)
           for _ in range(args.obs)]
    lambdas = {'surface': 2.0, 'reg_expr': 0.05, 'reg_global': 0.25, 'reg_unobserved': 10, 'reg_loc': 0.05, 'symm_dist': 5.0}
    schedule = {'lr': {200: 2, 400: 2}, 'symm_dist': {200: 10, 500: 9999}, 'reg_glob': {200: 3}, 'reg_loc': {500: 3}}

    def run(n):
        np.random.seed(0); torch.manual_seed(0)
        return inference_iterative_root_finding_joint(dec, dfn, obs, dict(lambdas), n_steps=n, schedule_cfg=schedule)

    run(3)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    z_ex, z_id, _ = run(args.steps)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    out = {'metric': 'joint_fit', 'iterations': args.steps, 'observations': args.obs, 'ms_per_iter': 1e3 * dt / args.steps,
           'iters_per_s': args.steps / dt, 'finite': bool(torch.isfinite(z_ex).all() and torch.isfinite(z_id).all())}
    print(json.dumps(out))
    if args.profile:
        from torch.profiler import profile, ProfilerActivity
        with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
            run(3)
            torch.cuda.synchronize()
        print(prof.key_averages().table(sort_by='cuda_time_total', row_limit=30, max_name_column_width=60))
        ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
        print('cuda kernel launches per iteration: %.0f   cuda time per iteration: %.2f ms'
              % (len(ev) / 3.0, sum(e.device_time for e in ev) / 3e3))


if __name__ == '__main__':
    main()

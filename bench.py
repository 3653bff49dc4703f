#!/usr/bin/env python
"""Benchmark of the NPHM hot path on B200 (contract: see the task statement / DESIGN.md section "Measurement").

A step = one pass of the hot path over one synthetic head: SDF of the 40-member ensemble on the 256^3 grid
(BASELINE.json configs[1]) followed by marching cubes on the resulting volume.
  value   whole-job SDF query points/s, inputs (latent, weights) resident in HBM, grid generated in-kernel
  e2e     same metric through the reference-facing drop-in API with HOST buffers: pinned latent -> H2D,
          get_logits(...) -> numpy volume (D2H), mesh_from_logits(numpy) -> mesh (H2D volume, D2H mesh)
  N > 1   weak scaling: every rank extracts its own head (independent latent), no data-path collective.

    python bench.py --gpus 1 --steps 5 --warmup 3
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference        # CPU arm: the oracle port on the host cores, bounded sample
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))

MINI = [-.55, -.5, -.95]
MAXI = [0.55, 0.75, 0.4]
FLOP_PER_POINT = 9.616e6          # dense reference formulation, SURVEY.md 8(d)
CHUNK = 25000                     # nbatch_points of scripts/fitting -sample (fitting_pointclouds.py:208)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='auto', choices=['auto', 'simt', 'tc', 'reference'])
    ap.add_argument('--res', type=int, default=256)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--cpu-sample-chunks', type=int, default=2)
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------ CPU arm
def cpu_reference_sample(res, n_chunks, with_mc=True):
    """Oracle port (numpy fp32, all host threads) on a bounded sample of the 256^3 workload: `n_chunks` chunks
    of 25 000 grid points through the ensemble + one full-volume marching cubes of a synthetic sphere volume
    (single-threaded C restatement).  Returns (points_per_s_estimate, detail dict)."""
    from conftest import load_golden, make_ensemble, sample_latent, sd_numpy
    from oracle import nphm_oracle as O
    threads = os.cpu_count() or 1
    dec = make_ensemble(0)
    p = O.EnsembleParams(sd_numpy(dec), load_golden('assets.npz')['anchors_39'])
    lat = sample_latent(1).numpy()
    total = res ** 3
    n = min(total, n_chunks * CHUNK)
    # sample = the middle of the grid (near the head, representative mix of members)
    first = max(0, total // 2 - n // 2)
    idx = np.arange(first, first + n)
    axes = [np.linspace(MINI[a], MAXI[a], res) for a in range(3)]
    pts = np.stack([axes[0][idx // (res * res)], axes[1][(idx // res) % res], axes[2][idx % res]], axis=1).astype(np.float32)
    t0 = time.perf_counter()
    O.get_logits(p, lat, pts, nbatch_points=CHUNK, threads=threads)
    t_sdf = time.perf_counter() - t0
    t_mc = 0.0
    if with_mc:
        ax = np.linspace(-0.5, 0.5, res, dtype=np.float32)
        vol = np.sqrt(ax[:, None, None] ** 2 + ax[None, :, None] ** 2 + ax[None, None, :] ** 2) - np.float32(0.37)
        t0 = time.perf_counter()
        O.marching_cubes(vol.astype(np.float32), 0.0)
        t_mc = time.perf_counter() - t0
    est_step = t_sdf * (total / n) + t_mc
    detail = {'sample_points': int(n), 't_sdf_sample_s': round(t_sdf, 3), 't_mc_full_s': round(t_mc, 3),
              'sdf_points_per_s': round(n / t_sdf, 1), 'est_step_s': round(est_step, 1), 'threads': threads}
    return total / est_step, detail


def workload_name(res):
    """One name for both arms of the benchmark (BASELINE.json configs[1] at res = 256)."""
    return ('single-head %d^3 grid SDF (40-member ensemble, seeded random weights) + marching cubes, random latent; '
            'one head per GPU' % res)


def run_reference_arm(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    vals = []
    detail = None
    for i in range(args.warmup + args.steps):
        v, detail = cpu_reference_sample(args.res, args.cpu_sample_chunks if i >= args.warmup else 1,
                                         with_mc=(i >= args.warmup))
        if i >= args.warmup:
            vals.append(v)
        if i == 0:
            args.warmup = min(args.warmup, 1)          # the CPU arm is slow: one warm-up pass is enough
    value = float(np.mean(vals))
    cores = os.cpu_count() or 1
    sample = ('%d chunks x %d grid points of the %d^3 grid through the oracle port (numpy fp32, %d threads) '
              '+ one full-volume C marching cubes; step time extrapolated to the full grid'
              % (args.cpu_sample_chunks, CHUNK, args.res, cores))
    line = {
        'impl': 'reference', 'metric': 'sdf_query_points_per_s', 'value': value, 'unit': 'points/s',
        'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': 1e3 * args.res ** 3 / value, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': workload_name(args.res), 'res': args.res, 'nbatch_points': CHUNK},
        'cpu_baseline': {'value': value, 'unit': 'points/s', 'cores': cores, 'kind': 'port', 'sample': sample,
                         'detail': detail},
        'e2e': {'value': value, 'unit': 'points/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []
        self.first = 0

    def start(self):
        if os.environ.get('NPHM_BENCH_NO_SAMPLER'):
            return
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.gpu), '--query-gpu=' + self.Q,
                                          '--format=csv,noheader,nounits', '-lms', os.environ.get('NPHM_BENCH_SAMPLE_MS', '200')],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def mark(self):
        """Start of the timed region: earlier samples (taken while nvidia-smi initialised, during warm-up) are dropped."""
        self.first = len(self.lines)

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for ln in self.lines[self.first:]:
            f = [x.strip() for x in ln.split(',')]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for name, val in zip(names, f[5:9]):
                if val.lower().startswith('active'):
                    reasons.add(name)
        if not sm:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['no samples']}
        return {'sm_mhz': float(np.median(sm)), 'sm_max_mhz': float(max(mx)), 'reasons': sorted(reasons),
                'power_w_max': max(power), 'samples': len(sm)}


# ------------------------------------------------------------------------------------------------ GPU arm
def main():
    args = parse()
    if args.impl == 'reference':
        run_reference_arm(args)
        return
    import torch
    import torch.distributed as dist
    from conftest import make_ensemble, sample_latent
    from nphm_b200 import _native
    from nphm_b200.models.reconstruction import get_logits
    from nphm_b200.utils.reconstruction import create_grid_points_from_bounds, mesh_from_logits

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    os.environ['NPHM_B200_IMPL'] = args.impl

    res = args.res
    total = res ** 3
    dec = make_ensemble(0, device=dev).eval()
    eng = dec.engine()
    lat = sample_latent(1 + rank).to(dev)                       # every rank its own random head
    volume = torch.empty(total, device=dev, dtype=torch.float32)
    flush = torch.empty(64 * 1024 * 1024, device=dev, dtype=torch.float32)      # 256 MB > 126 MB L2
    launches = {'n': 0}

    # kernels of this library per step (ncu launch list: profiles/r01_launches_bench256_final.txt): grid axes, anchors, folded
    # constants, [tensor-core records], ensemble | marching cubes: classify, scan, vertices, triangles
    per_step_launches = 8 + (1 if args.impl in ('auto', 'tc') else 0)

    def step_device():
        """grid SDF (in-kernel grid) + marching cubes, everything resident on the device."""
        eng.query_grid(lat, MINI, MAXI, res, 0, total, quirk_period=CHUNK, out=volume)
        v, t = _native.marching_cubes_device(volume.view(res, res, res), 0.0, negate=True)
        launches['n'] += per_step_launches
        return v, t

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing --------------------------------------------------------------------------
    sampler = ClockSampler(local)
    n_warm = max(args.warmup, 3)
    for i in range(n_warm):
        if i == n_warm - 1:
            sampler.start()             # nvidia-smi takes driver locks while it initialises: let that happen in warm-up
        step_device()
        flush.zero_()
    import gc
    gc.collect()
    gc.disable()                        # no collector pauses inside the timed regions (host sits between MC passes)
    barrier()
    sampler.mark()
    launches['n'] = 0
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    kev = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(args.steps)]
    ev[0].record()
    n_tris = 0
    for i in range(args.steps):
        kev[i][0].record()
        eng.query_grid(lat, MINI, MAXI, res, 0, total, quirk_period=CHUNK, out=volume)
        kev[i][1].record()
        v, t = _native.marching_cubes_device(volume.view(res, res, res), 0.0, negate=True)
        kev[i][2].record()
        launches['n'] += per_step_launches
        n_tris = t.shape[0]
        del v, t                        # like the warm-up: the mesh buffers go back to torch's caching allocator (no cudaMalloc
                                        # for a second generation of buffers inside the timed region)
        flush.zero_()                   # L2 flush between timed iterations (inside the timed region)
        kev[i][3].record()
    ev[1].record()
    barrier()
    clocks = sampler.stop()
    ms_total = ev[0].elapsed_time(ev[1])
    if os.environ.get('NPHM_BENCH_DEBUG'):
        print('per-step ms (sdf, mc, flush):', [[round(k[j].elapsed_time(k[j + 1]), 3) for j in range(3)] for k in kev], file=sys.stderr)
    sdf_ms = float(np.mean([k[0].elapsed_time(k[1]) for k in kev]))   # prep kernels + the ensemble kernel
    mc_ms = float(np.mean([k[1].elapsed_time(k[2]) for k in kev]))    # marching cubes (4 kernels + count readback)
    flush_ms = float(np.mean([k[2].elapsed_time(k[3]) for k in kev]))
    t = torch.tensor([ms_total, sdf_ms, mc_ms, flush_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total, sdf_ms, mc_ms, flush_ms = t.tolist()
    ms_per_step = ms_total / args.steps
    value = world * total / (ms_per_step * 1e-3)

    # ---- opt-in pruned kernel (reported separately; NOT the dense reference computation) ---------------------
    pruned = None
    if world == 1 and args.impl in ('auto', 'tc'):
        tau = 1e-8
        eng.set_prune_threshold(tau)
        ref_vol = volume.clone()
        pv = torch.empty_like(volume)
        for _ in range(2):
            eng.query_grid(lat, MINI, MAXI, res, 0, total, quirk_period=CHUNK, out=pv, impl='tc_pruned')
        torch.cuda.synchronize()
        p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        p0.record()
        for _ in range(args.steps):
            eng.query_grid(lat, MINI, MAXI, res, 0, total, quirk_period=CHUNK, out=pv, impl='tc_pruned')
        p1.record()
        torch.cuda.synchronize()
        pms = p0.elapsed_time(p1) / args.steps
        pruned = {'tau': tau, 'sdf_ms': pms, 'value': total / (pms * 1e-3), 'unit': 'points/s',
                  'max_abs_diff_vs_dense': float((pv - ref_vol).abs().max().item()),
                  'note': 'NPHM_IMPL_TC_PRUNED: members with normalised blend weight < tau on a whole 8x4x4 tile are skipped'}

    # ---- end to end through the drop-in API with host buffers ---------------------------------------------
    grid_points = torch.from_numpy(create_grid_points_from_bounds(MINI, MAXI, res)).to(dev, dtype=torch.float)
    grid_points = grid_points.reshape(1, -1, 3)                 # uploaded once, like fitting_pointclouds.py:166-170
    lat_host = lat.cpu().pin_memory()
    h2d = d2h = 0

    def step_e2e():
        nonlocal h2d, d2h
        z = lat_host.to(dev, non_blocking=True)                                  # H2D: latent
        logits = get_logits(dec, z, grid_points, nbatch_points=CHUNK)            # D2H: volume (numpy)
        mesh = mesh_from_logits(logits, MINI, MAXI, res)                         # H2D volume, GPU MC, D2H mesh
        h2d = lat_host.numel() * 4 + logits.nbytes
        d2h = logits.nbytes + np.asarray(mesh.vertices).nbytes + np.asarray(mesh.faces).nbytes
        return mesh

    for _ in range(2):
        step_e2e()
    barrier()
    t0 = time.perf_counter()
    e2e_steps = max(2, min(args.steps, 5))
    for _ in range(e2e_steps):
        step_e2e()
    barrier()
    e2e_s = torch.tensor([(time.perf_counter() - t0) / e2e_steps], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_value = world * total / e2e_s.item()
    gc.enable()

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
        except Exception:
            pass
        peak_tf = peaks.get('bf16_tflops_sustained', None)
        peak_src = 'measured (MEASURED_PEAKS.json bf16_tflops_sustained)'
        if peak_tf is None:
            peak_tf, peak_src = 1590.0, 'fallback (B200_PROFILING.md)'
        achieved_tf = FLOP_PER_POINT * total / (sdf_ms * 1e-3) / 1e12
        # DRAM bytes of one launch of the dominant kernel, from the committed ncu --set full capture of this workload
        traffic, traffic_src = None, 'no ncu capture recorded for this resolution'
        try:
            with open(os.path.join(ROOT, 'profiles', 'traffic.json')) as f:
                rec = json.load(f).get(str(res))
            if rec:
                traffic = rec['dram_bytes_read'] + rec['dram_bytes_write']
                traffic_src = 'bytes, ' + rec['source'] + ' (profiles/traffic.json)'
        except (OSError, ValueError, KeyError):
            pass
        line = {
            'metric': 'sdf_query_points_per_s', 'value': value, 'unit': 'points/s', 'n_gpus': world,
            'steps': args.steps, 'warmup': max(args.warmup, 3), 'ms_per_step': ms_per_step,
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': workload_name(res),
                       'res': res, 'nbatch_points': CHUNK, 'impl': args.impl, 'l2': 'flushed between iterations (256 MB write)',
                       'triangles': int(n_tris)},
            'meshes_per_s': world / (ms_per_step * 1e-3),
            'sdf_ms': sdf_ms, 'mc_ms': mc_ms, 'l2_flush_ms': flush_ms,
            'gpu_launches': launches['n'],
            'clocks': clocks,
            'e2e': {'value': e2e_value, 'unit': 'points/s', 'h2d_bytes_per_step': int(h2d), 'd2h_bytes_per_step': int(d2h),
                    'api': 'get_logits + mesh_from_logits (drop-in, numpy in/out)'},
            'roofline': {'bound': 'tensor', 'achieved': achieved_tf, 'peak': peak_tf, 'unit': 'TFLOP/s',
                         'frac': achieved_tf / peak_tf, 'traffic': traffic,
                         'note': 'dominant kernel = fused ensemble SDF query; algorithmic 9.616 MFLOP/point (dense '
                                 'reference formulation) / CUDA-event time of the query; peak: ' + peak_src +
                                 '; traffic: ' + traffic_src},
        }
        if pruned is not None:
            line['pruned_opt_in'] = pruned
        if not args.no_cpu_baseline and world == 1:
            v, detail = cpu_reference_sample(res, args.cpu_sample_chunks)
            line['cpu_baseline'] = {'value': v, 'unit': 'points/s', 'cores': os.cpu_count() or 1, 'kind': 'port',
                                    'sample': '%d x %d grid points through the oracle port (numpy fp32, all threads) + one '
                                              'full-volume C marching cubes, extrapolated to %d^3' % (args.cpu_sample_chunks, CHUNK, res),
                                    'detail': detail}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()

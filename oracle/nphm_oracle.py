"""ORACLE - test infrastructure only.

Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py`` may import this module; nothing under ``nphm_b200/`` does.  It is a plain numpy (fp32)
restatement of the reference's algorithm for the hot path, each function citing the reference lines it
follows (paths relative to /root/reference).  It is written in the reference's *dense* formulation
(inputs concatenated, no constant folding) so that it is independent of the CUDA kernels' algebra.

Pinned against the reference itself: ``tests/golden/make_golden.py`` imports the reference's PyTorch
modules in the build container and stores their outputs; ``tests/test_oracle.py`` checks this module
against those vectors.  The marching-cubes part lives in ``mc_oracle.c`` (parity unpinned, see there).
"""
from __future__ import annotations

import ctypes
import os
from concurrent.futures import ThreadPoolExecutor

import numpy as np

F32 = np.float32


# --------------------------------------------------------------------------------------------------
# elementary pieces
# --------------------------------------------------------------------------------------------------
def softplus100(x: np.ndarray) -> np.ndarray:
    """``nn.Softplus(beta=100)`` with torch's default threshold 20:
    x if 100*x > 20 else log1p(exp(100*x))/100.   (EnsembledDeepSDF.py:99, deepSDF.py:56-57)"""
    x = x.astype(F32, copy=False)
    bx = x * F32(100.0)
    with np.errstate(over='ignore'):
        soft = np.log1p(np.exp(bx)) / F32(100.0)
    return np.where(bx > F32(20.0), x, soft).astype(F32)


def member_weight_set(n_members: int, n_symm: int) -> np.ndarray:
    """EnsembledDeepSDF.py:43-45: members (2i, 2i+1), i < n_symm share set i; the rest follow."""
    k = np.arange(n_members)
    return np.where(k < 2 * n_symm, k // 2, k - n_symm)


def linspace_grid(mini, maxi, res: int) -> np.ndarray:
    """utils/reconstruction.py:5-20 (scale=None): (res^3, 3) float64, z fastest."""
    x = np.linspace(mini[0], maxi[0], res)
    y = np.linspace(mini[1], maxi[1], res)
    z = np.linspace(mini[2], maxi[2], res)
    X, Y, Z = np.meshgrid(x, y, z, indexing='ij')
    return np.column_stack((X.reshape(-1), Y.reshape(-1), Z.reshape(-1)))


# --------------------------------------------------------------------------------------------------
# ensemble (identity SDF)
# --------------------------------------------------------------------------------------------------
class EnsembleParams:
    """Plain-numpy view of a ``FastEnsembleDeepSDFMirrored.state_dict()`` + its mean anchors."""

    def __init__(self, state_dict, mean_anchors, lat_dim_glob=64, lat_dim_loc=32, n_loc=39, n_symm=16):
        g = lambda k: np.asarray(state_dict[k], dtype=F32)
        self.n_lin = len([k for k in state_dict if k.startswith('ensembled_deep_sdf.lin') and k.endswith('weight')])
        self.W = [g('ensembled_deep_sdf.lin%d.weight' % i) for i in range(self.n_lin)]   # (sets, out, in)
        self.b = [g('ensembled_deep_sdf.lin%d.bias' % i) for i in range(self.n_lin)]
        self.pos_W = [g('mlp_pos.%d.weight' % i) for i in (0, 2, 4)]
        self.pos_b = [g('mlp_pos.%d.bias' % i) for i in (0, 2, 4)]
        self.mean_anchors = np.asarray(mean_anchors, dtype=F32).reshape(n_loc, 3)
        self.G, self.L, self.K, self.n_symm = lat_dim_glob, lat_dim_loc, n_loc, n_symm
        self.A = n_loc + 1
        nlayers = self.n_lin - 1
        self.skip = nlayers // 2
        self.ws = member_weight_set(self.A, n_symm)


def predict_anchors(p: EnsembleParams, latent: np.ndarray) -> np.ndarray:
    """EnsembledDeepSDF.py:228-229: anchors = mlp_pos(z_glob).view(K,3) + mean anchors."""
    h = latent[:p.G].astype(F32)
    h = np.maximum(p.pos_W[0] @ h + p.pos_b[0], 0)
    h = np.maximum(p.pos_W[1] @ h + p.pos_b[1], 0)
    h = p.pos_W[2] @ h + p.pos_b[2]
    return (h.reshape(p.K, 3) + p.mean_anchors).astype(F32)


def _member_forward(p: EnsembleParams, k: int, inp_k: np.ndarray) -> np.ndarray:
    """One member's DeepSDF stack on its (N, 99) input: EnsembledDeepSDF.py:101-126 with weight set ws[k]."""
    s = p.ws[k]
    h = inp_k
    inv = F32(np.sqrt(2))
    for layer in range(p.n_lin):
        if layer == p.skip:
            h = (np.concatenate([h, inp_k], axis=1) / inv).astype(F32)             # :115-116
        h = h @ p.W[layer][s].T + p.b[layer][s]                                    # :48,:54
        if layer < p.n_lin - 1:
            h = softplus100(h)                                                     # :120-121
    return h[:, 0]


def _ensemble_members(p: EnsembleParams, xyz: np.ndarray, latent: np.ndarray, anchors: np.ndarray,
                      pool=None) -> np.ndarray:
    """s_k(x) for all members: EnsembledDeepSDF.py:240-257 + :101-126.  Returns (N, A)."""
    N = xyz.shape[0]
    G, L, A = p.G, p.L, p.A
    origin = np.concatenate([anchors, np.zeros((1, 3), F32)], axis=0)             # :240-241
    coords = xyz[:, None, :] - origin[None, :, :]                                  # N x A x 3
    coords[:, 1:2 * p.n_symm:2, 0] *= F32(-1)                                      # :244
    z_loc = latent[G:].reshape(A, L)                                               # :248
    cond = np.concatenate([np.broadcast_to(latent[:G], (A, G)), z_loc], axis=1)    # :247-252
    out = np.empty((A, N), F32)

    def run(k):
        inp_k = np.concatenate([coords[:, k, :], np.broadcast_to(cond[k], (N, G + L))], axis=1).astype(F32)
        out[k] = _member_forward(p, k, inp_k)

    if pool is not None:
        list(pool.map(run, range(A)))
    else:
        for k in range(A):
            run(k)
    return out.T


def blend(xyz: np.ndarray, anchors: np.ndarray, s: np.ndarray) -> np.ndarray:
    """sample_point_feature(..., background=True, var=0.01): EnsembledDeepSDF.py:129-150."""
    diff = anchors[None, :, :] - xyz[:, None, :]
    nrm = np.sqrt((diff * diff).sum(axis=2, dtype=F32)).astype(F32)
    dist = -((nrm + F32(10e-6)) ** 2)
    dist = np.concatenate([dist, np.full((xyz.shape[0], 1), -0.2, F32)], axis=1).astype(F32)
    w = np.exp(dist / F32(0.1 ** 2)).astype(F32)
    w = w / (w.sum(axis=1, keepdims=True, dtype=F32) + F32(1e-6))
    return (w * s).sum(axis=1, dtype=F32).astype(F32)


def ensemble_forward(p: EnsembleParams, xyz: np.ndarray, latent: np.ndarray, eval_mode: bool = True,
                     chunk: int = 8192, threads: int = 1):
    """``FastEnsembleDeepSDFMirrored.forward(xyz[None], latent[None,None], None)`` for one call:
    in eval mode the LAST point of the call gets s_k = 1 for every member (:260-261).
    ``threads`` > 1 evaluates the members of a chunk concurrently (one BLAS thread each).
    Returns (sdf (N,), anchors (K,3))."""
    xyz = np.ascontiguousarray(xyz, dtype=F32)
    latent = np.asarray(latent, dtype=F32)
    anchors = predict_anchors(p, latent)
    N = xyz.shape[0]
    out = np.empty(N, F32)

    def run_all(pool):
        for lo in range(0, N, chunk):
            hi = min(N, lo + chunk)
            s = _ensemble_members(p, xyz[lo:hi], latent, anchors, pool)
            if eval_mode and hi == N:
                s[-1, :] = 1
            out[lo:hi] = blend(xyz[lo:hi], anchors, s)

    if threads > 1:
        try:
            from threadpoolctl import threadpool_limits
        except ImportError:                                     # pragma: no cover
            threadpool_limits = None
        with ThreadPoolExecutor(threads) as pool:
            if threadpool_limits is not None:
                with threadpool_limits(limits=1):
                    run_all(pool)
            else:
                run_all(pool)
    else:
        run_all(None)
    return out, anchors


def get_logits(p: EnsembleParams, latent, grid_points: np.ndarray, nbatch_points: int = 100000,
               eval_mode: bool = True, threads: int = 1) -> np.ndarray:
    """models/reconstruction.py:6-25: the grid is split in chunks of ``nbatch_points`` and every chunk is
    one decoder call - so in eval mode the last point of EVERY chunk carries the :260-261 quirk."""
    pts = np.asarray(grid_points, dtype=F32).reshape(-1, 3)
    outs = []
    for lo in range(0, pts.shape[0], nbatch_points):
        s, _ = ensemble_forward(p, pts[lo:lo + nbatch_points], latent, eval_mode=eval_mode, threads=threads)
        outs.append(s)
    return np.concatenate(outs)


# --------------------------------------------------------------------------------------------------
# DeepSDF backbone / deformation network
# --------------------------------------------------------------------------------------------------
class MlpParams:
    """Plain-numpy view of a ``DeepSDF.state_dict()`` (prefix '' or 'defDeepSDF.')."""

    def __init__(self, state_dict, prefix=''):
        keys = [k for k in state_dict if k.startswith(prefix + 'lin') and k.endswith('.weight')]
        self.n_lin = len(keys)
        self.W = [np.asarray(state_dict['%slin%d.weight' % (prefix, i)], dtype=F32) for i in range(self.n_lin)]
        self.b = [np.asarray(state_dict['%slin%d.bias' % (prefix, i)], dtype=F32) for i in range(self.n_lin)]
        self.skip = (self.n_lin - 1) // 2


def mlp_forward(m: MlpParams, xyz: np.ndarray, cond: np.ndarray) -> np.ndarray:
    """DeepSDF.forward with a per-call constant latent: deepSDF.py:64-89 (no positional encoding)."""
    xyz = np.asarray(xyz, dtype=F32)
    inp = np.concatenate([xyz, np.broadcast_to(np.asarray(cond, F32), (xyz.shape[0], cond.shape[-1]))], axis=1)
    h = inp
    for layer in range(m.n_lin):
        if layer == m.skip:
            h = (np.concatenate([h, inp], axis=1) / F32(np.sqrt(2))).astype(F32)
        h = h @ m.W[layer].T + m.b[layer]
        if layer < m.n_lin - 1:
            h = softplus100(h)
    return h.astype(F32)


def deformation_forward(state_dict, xyz, z_id, z_ex, anchors) -> np.ndarray:
    """DeformationNetwork.forward, mode 'compress', eval mode: deepSDF.py:212-239.
    cond = [Linear_1461->32([z_id | anchors.flatten()]) | z_ex]; returns offsets (N,3)."""
    Wc = np.asarray(state_dict['compressor.0.weight'], F32)
    bc = np.asarray(state_dict['compressor.0.bias'], F32)
    first = np.concatenate([np.asarray(z_id, F32), np.asarray(anchors, F32).reshape(-1)])
    compressed = Wc @ first + bc
    cond = np.concatenate([compressed, np.asarray(z_ex, F32)])
    m = MlpParams(state_dict, prefix='defDeepSDF.')
    return mlp_forward(m, xyz, cond)[:, :3]


# --------------------------------------------------------------------------------------------------
# Broyden correspondence search (iterative_root_finding.py:5-71 `broyden`, residual of `search` :142-147)
# --------------------------------------------------------------------------------------------------
def broyden_search(field, obs, x_init, J_inv_init, max_steps=15, cvg_thresh=1e-6, dvg_thresh=0.2, eps=1e-6):
    """Roots of ``x + field(x) - obs`` per row.  ``field``: (N,3) -> (N,3) fp32 (rows independent); obs, x_init (N,3);
    J_inv_init (N,3,3).  Returns (x, diff, valid): like the reference, ``x`` is where each sample stopped (its `x_opt`
    aliases `x`, :34), ``diff`` the smallest residual norm seen, ``valid = diff < cvg_thresh``."""
    obs = np.asarray(obs, F32)
    x = np.array(x_init, F32)
    J = np.array(J_inv_init, F32)
    gx = ((field(x) + x) - obs).astype(F32)
    upd = -np.einsum('nij,nj->ni', J, gx).astype(F32)
    best = np.sqrt((gx * gx).sum(-1)).astype(F32)
    act = np.ones(x.shape[0], bool)
    for _ in range(max_steps):
        dx = upd                                                      # rows of the currently active samples
        x[act] = x[act] + dx
        gnew = ((field(x) + x) - obs).astype(F32)[act]
        dg = (gnew - gx[act]).astype(F32)
        gx[act] = gx[act] + dg
        nrm = np.sqrt((gx * gx).sum(-1)).astype(F32)
        better = nrm < best
        best[better] = nrm[better]
        new_act = (best > cvg_thresh) & (nrm < dvg_thresh)
        if not new_act.any():
            break
        sel = new_act[act]                                            # active sets are nested
        dx, dg = dx[sel], dg[sel]
        Ja = J[new_act]
        vT = np.einsum('ni,nij->nj', dx, Ja).astype(F32)
        a = (dx - np.einsum('nij,nj->ni', Ja, dg)).astype(F32)
        b = (vT * dg).sum(-1).astype(F32)
        b = np.where(b >= 0, b + F32(eps), b - F32(eps)).astype(F32)
        Ja = (Ja + (a / b[:, None])[:, :, None] * vT[:, None, :]).astype(F32)
        J[new_act] = Ja
        upd = -np.einsum('nij,nj->ni', Ja, gx[new_act]).astype(F32)
        act = new_act
    return x, best, best < cvg_thresh


# --------------------------------------------------------------------------------------------------
# Adam (fitting.py:35, torch.optim.Adam defaults, torch 2.11 single-tensor update order)
# --------------------------------------------------------------------------------------------------
def adam_step(param, grad, m, v, step, lr, beta1=0.9, beta2=0.999, eps=1e-8):
    """One ``torch.optim.Adam`` step (no weight decay / amsgrad), fp32 state, python-float scalars.
    Returns (param, m, v) updated; ``step`` is the 1-based step count."""
    param = param.astype(F32); grad = grad.astype(F32)
    m = (m + (grad - m) * F32(1 - beta1)).astype(F32)                 # exp_avg.lerp_(grad, 1-beta1)
    v = (v * F32(beta2) + F32(1 - beta2) * grad * grad).astype(F32)   # mul_(beta2).addcmul_(g, g, 1-beta2)
    bc1 = 1 - beta1 ** step
    bc2 = 1 - beta2 ** step
    step_size = lr / bc1
    denom = (np.sqrt(v) / F32(np.sqrt(bc2)) + F32(eps)).astype(F32)
    param = (param - F32(step_size) * (m / denom)).astype(F32)
    return param, m, v


# --------------------------------------------------------------------------------------------------
# marching cubes (C restatement, built by __graft_entry__.build() / oracle/Makefile)
# --------------------------------------------------------------------------------------------------
_MC = None


def _mc_lib():
    global _MC
    if _MC is None:
        here = os.path.dirname(os.path.abspath(__file__))
        path = os.path.join(here, '_build', 'libmc_oracle.so')
        if not os.path.exists(path):
            raise RuntimeError('oracle marching cubes not built: run `make -C oracle` '
                               '(or __graft_entry__.build())')
        lib = ctypes.CDLL(path)
        lib.mc_oracle.restype = ctypes.c_int
        lib.mc_oracle.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_double,
                                  ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p,
                                  ctypes.POINTER(ctypes.c_longlong), ctypes.POINTER(ctypes.c_longlong)]
        _MC = lib
    return _MC


def marching_cubes(volume: np.ndarray, isovalue: float = 0.0, negate: bool = False):
    """``mcubes.marching_cubes(volume, isovalue)`` restated: returns (verts (V,3) f64 in index units,
    tris (T,3) uint64)."""
    lib = _mc_lib()
    vol = np.ascontiguousarray(volume, dtype=F32)
    nx, ny, nz = vol.shape
    nv, nt = ctypes.c_longlong(0), ctypes.c_longlong(0)
    rc = lib.mc_oracle(vol.ctypes.data, nx, ny, nz, float(isovalue), int(negate), None, None,
                       ctypes.byref(nv), ctypes.byref(nt))
    assert rc == 0
    verts = np.empty((nv.value, 3), np.float64)
    tris = np.empty((nt.value, 3), np.uint64)
    rc = lib.mc_oracle(vol.ctypes.data, nx, ny, nz, float(isovalue), int(negate),
                       verts.ctypes.data, tris.ctypes.data, ctypes.byref(nv), ctypes.byref(nt))
    assert rc == 0
    return verts, tris


def mesh_from_logits(logits: np.ndarray, mini, maxi, resolution: int):
    """utils/reconstruction.py:22-37 without the trimesh wrapper: negates ``logits`` IN PLACE like the
    reference, runs MC at 0, rescales: verts*step + mini (float64)."""
    logits = np.reshape(logits, (resolution,) * 3)
    logits *= -1
    verts, tris = marching_cubes(logits, 0.0)
    step = (np.array(maxi) - np.array(mini)) / (resolution - 1)
    verts = verts * np.expand_dims(step, axis=0)
    verts += [mini[0], mini[1], mini[2]]
    return verts, tris


# --------------------------------------------------------------------------------------------------
# identity-space fitting: loss and latent gradient (fitting.py:197-279), dense formulation, manual backprop
# --------------------------------------------------------------------------------------------------
def _sigmoid100(a: np.ndarray) -> np.ndarray:
    """d softplus_100 / da as torch's backward computes it: 1 beyond the threshold, else e/(e+1)."""
    ba = a.astype(F32) * F32(100.0)
    with np.errstate(over='ignore'):
        e = np.exp(np.minimum(ba, F32(30.0)))
    return np.where(ba > F32(20.0), F32(1.0), e / (e + F32(1.0))).astype(F32)


def fit_identity_loss_grad(p: EnsembleParams, points: np.ndarray, z: np.ndarray, lambdas: dict, clamp: float):
    """One evaluation of the loss of ``inference_identity_space`` and its gradient w.r.t. the latent ``z``.

    points: (N,3) sampled observation points (already stacked), z: (lat_dim,).  Loss terms
    (fitting.py:239-268): surface = mean(|sdf|[|sdf| < clamp]), reg_global = |z[:64]|^2, reg_loc = |z[64:]|^2,
    reg_unobserved = sum_{k in 30,31,39} |z_k|^2, symm_dist = mean_i |z_2i - z_2i+1|.
    Returns (terms dict, grad (lat_dim,), n_kept)."""
    x = np.ascontiguousarray(points, dtype=F32)
    z = np.asarray(z, dtype=F32)
    N = x.shape[0]
    G, L, A, K = p.G, p.L, p.A, p.K
    r2 = F32(np.sqrt(2))
    # ---- anchors = mlp_pos(z_g) + mean  (forward with ReLU masks kept)
    zg = z[:G]
    a0 = p.pos_W[0] @ zg + p.pos_b[0]; h0 = np.maximum(a0, 0)
    a1 = p.pos_W[1] @ h0 + p.pos_b[1]; h1 = np.maximum(a1, 0)
    anchors = ((p.pos_W[2] @ h1 + p.pos_b[2]).reshape(K, 3) + p.mean_anchors).astype(F32)
    # ---- per-member forward, keeping pre-activations
    origin = np.concatenate([anchors, np.zeros((1, 3), F32)], axis=0)
    coords = x[:, None, :] - origin[None, :, :]
    flip = np.ones((A, 3), F32); flip[1:2 * p.n_symm:2, 0] = -1
    coords = coords * flip[None]
    z_loc = z[G:].reshape(A, L)
    cond = np.concatenate([np.broadcast_to(zg, (A, G)), z_loc], axis=1)
    saved = []
    s = np.empty((N, A), F32)
    for k in range(A):
        ws = p.ws[k]
        inp = np.concatenate([coords[:, k, :], np.broadcast_to(cond[k], (N, G + L))], axis=1).astype(F32)
        pre = []
        h = inp
        for layer in range(p.n_lin):
            if layer == p.skip:
                h = (np.concatenate([h, inp], axis=1) / r2).astype(F32)
            a = h @ p.W[layer][ws].T + p.b[layer][ws]
            pre.append((h, a))
            h = softplus100(a) if layer < p.n_lin - 1 else a
        s[:, k] = h[:, 0]
        saved.append((inp, pre))
    # ---- blend
    diff = anchors[None] - x[:, None, :]
    r = np.sqrt((diff * diff).sum(axis=2)).astype(F32)
    nrm = r + F32(10e-6)
    dist = np.concatenate([-(nrm ** 2), np.full((N, 1), -0.2, F32)], axis=1)
    w = np.exp(dist / F32(0.1 ** 2)).astype(F32)
    S = w.sum(axis=1) + F32(1e-6)
    out = (w * s).sum(axis=1) / S
    l = np.abs(out)
    kept = l < F32(clamp)
    n_kept = int(kept.sum())
    terms = {'surface': (l[kept].mean() if n_kept else np.float32('nan'))}
    # ---- backward: surface term
    g_out = np.where(kept, np.sign(out), 0).astype(F32) * F32(lambdas.get('surface', 0.0)) / F32(max(n_kept, 1))
    if n_kept == 0:
        g_out = g_out * np.float32('nan')
    g_s = g_out[:, None] * w / S[:, None]
    g_w = g_out[:, None] * (s - out[:, None]) / S[:, None]
    g_dist = g_w * w / F32(0.1 ** 2)
    with np.errstate(divide='ignore', invalid='ignore'):
        coef = np.where(r > 0, g_dist[:, :K] * F32(-2.0) * nrm / r, 0).astype(F32)
    g_anchor = (coef[:, :, None] * diff).sum(axis=0)                       # blend path, (K,3)
    grad = np.zeros_like(z)
    for k in range(A):
        ws = p.ws[k]
        inp, pre = saved[k]
        g = g_s[:, k:k + 1] * p.W[p.n_lin - 1][ws]                          # dL/dh3
        g_inp = np.zeros_like(inp)
        for layer in range(p.n_lin - 2, -1, -1):
            h_in, a = pre[layer]
            g_a = g * _sigmoid100(a)
            g = g_a @ p.W[layer][ws]
            if layer == p.skip:
                g = g / r2
                n_h = g.shape[1] - inp.shape[1]
                g_inp += g[:, n_h:]
                g = g[:, :n_h]
        g_inp += g
        g_c = g_inp[:, :3].sum(axis=0)
        g_u = g_inp[:, 3:].sum(axis=0)
        grad[:G] += g_u[:G]
        grad[G + k * L:G + (k + 1) * L] += g_u[G:]
        if k < K:
            g_anchor[k] += -g_c * flip[k]                                  # c = (x - a) * flip
    # ---- anchors -> z_glob through mlp_pos
    g_o = g_anchor.reshape(-1).astype(F32)
    g_h1 = (p.pos_W[2].T @ g_o) * (a1 > 0)
    g_h0 = (p.pos_W[1].T @ g_h1) * (a0 > 0)
    grad[:G] += p.pos_W[0].T @ g_h0
    # ---- regularisers
    terms['reg_global'] = float((z[:G] ** 2).sum())
    terms['reg_loc'] = float((z[G:] ** 2).sum())
    grad[:G] += F32(lambdas.get('reg_global', 0.0)) * 2 * z[:G]
    grad[G:] += F32(lambdas.get('reg_loc', 0.0)) * 2 * z[G:]
    ru = 0.0
    for k in (30, 31, 39):
        zk = z[G + k * L:G + (k + 1) * L]
        ru += float((zk ** 2).sum())
        grad[G + k * L:G + (k + 1) * L] += F32(lambdas.get('reg_unobserved', 0.0)) * 2 * zk
    terms['reg_unobserved'] = ru
    pairs = z[G:G + 2 * p.n_symm * L].reshape(p.n_symm, 2, L)
    dvec = pairs[:, 0] - pairs[:, 1]
    dn = np.sqrt((dvec ** 2).sum(axis=1))
    terms['symm_dist'] = float(dn.mean()) if p.n_symm else 0.0
    with np.errstate(divide='ignore', invalid='ignore'):
        gd = np.where(dn[:, None] > 0, dvec / dn[:, None], 0) / max(p.n_symm, 1) * F32(lambdas.get('symm_dist', 0.0))
    gpairs = np.stack([gd, -gd], axis=1).reshape(-1)
    grad[G:G + 2 * p.n_symm * L] += gpairs.astype(F32)
    return terms, grad.astype(F32), n_kept

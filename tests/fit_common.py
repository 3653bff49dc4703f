"""Shared helpers for the fitting tests: replay the reference's sampling / schedule for the golden run
(tests/golden/make_golden.py, section 'identity-space fitting trajectory')."""
import numpy as np
import torch

from conftest import load_golden


def golden_fit_setup():
    g = load_golden('fit_identity.npz')
    obs = [torch.from_numpy(o) for o in g['obs']]
    lambdas = {'surface': 2.0, 'reg_global': 0.25, 'reg_unobserved': 10, 'reg_loc': 0.05, 'symm_dist': 5.0}
    schedule = {'lr': {200: 2, 400: 2, 600: 2, 800: 2}, 'symm_dist': {200: 10, 500: 9999},
                'reg_glob': {200: 3, 600: 10}, 'reg_loc': {500: 3, 600: 10}}
    return g, obs, lambdas, schedule


def replay_iterations(n_iter, step_scale=0.01):
    """Yields (j, points (5*n,3) numpy, lambdas snapshot, clamp, lr) exactly as the reference loop saw them."""
    from nphm_b200.models.fitting import _apply_schedule, _clamp_for_iteration, _sample_observations
    g, obs, lambdas, schedule = golden_fit_setup()
    np.random.seed(0)
    torch.manual_seed(0)
    lr = 0.01
    for j in range(n_iter):
        lr = _apply_schedule(j, step_scale, schedule, lambdas, lr)
        pts, _ = _sample_observations(obs)
        yield j, pts.reshape(-1, 3).numpy().copy(), dict(lambdas), _clamp_for_iteration(j, step_scale), lr

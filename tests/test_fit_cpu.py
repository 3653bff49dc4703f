"""CPU suite, part 3: the fitting path - oracle gradient pinned to the reference's autograd (golden), host-side
schedule / sampling logic, and the autograd fallback of the drop-in fitter."""
import numpy as np
import torch

from conftest import load_golden, make_ensemble, sd_numpy
from fit_common import golden_fit_setup, replay_iterations
from oracle import nphm_oracle as O


def _rel(a, b):
    return float(np.abs(a - b).max() / (np.abs(b).max() + 1e-30))


def test_replayed_schedule_matches_reference():
    g, _, lambdas, _ = golden_fit_setup()
    its = list(replay_iterations(12))
    lrs = np.array([it[4] for it in its])
    assert np.allclose(lrs, g['lrs'], rtol=0, atol=0)                      # lr halved at j = 2, 4, 6, 8
    final = its[-1][2]
    assert np.allclose([final[k] for k in sorted(final)], g['lambdas_final'])
    assert [it[3] for it in its][:7] == [0.1, 0.1, 0.1, 0.05, 0.05, 0.05, 0.0075]


def test_oracle_fit_gradient_matches_reference_autograd():
    g, _, _, _ = golden_fit_setup()
    p = O.EnsembleParams(sd_numpy(make_ensemble(0)), load_golden('assets.npz')['anchors_39'])
    for j, pts, lam, clamp, lr in replay_iterations(8):
        if j not in (0, 1, 3, 7):
            continue
        terms, grad, n_kept = O.fit_identity_loss_grad(p, pts, g['z_before'][j], lam, clamp)
        ref = g['grads'][j]
        assert n_kept > 0
        assert _rel(grad, ref) < 2e-4, (j, _rel(grad, ref))


def test_oracle_adam_replays_reference_trajectory():
    """Feeding the reference's own gradients through the oracle's Adam reproduces its latent trajectory."""
    g, _, _, _ = golden_fit_setup()
    z = np.zeros(1344, np.float32); m = np.zeros_like(z); v = np.zeros_like(z)
    for j in range(12):
        assert np.abs(z - g['z_before'][j]).max() < 1e-6
        z, m, v = O.adam_step(z, g['grads'][j], m, v, j + 1, float(g['lrs'][j]))
    assert np.abs(z - g['z_final']).max() < 1e-6


def test_autograd_fallback_follows_reference():
    """3 iterations of the drop-in fitter on CPU (composite modules + torch autograd) against the reference's latents."""
    from nphm_b200.models.fitting import inference_identity_space
    g, obs, lambdas, schedule = golden_fit_setup()
    dec = make_ensemble(0).train()
    np.random.seed(0)
    torch.manual_seed(0)
    z, anchors = inference_identity_space(dec, obs, lambdas, n_steps=300, schedule_cfg=schedule, step_scale=0.01)
    assert z.shape == (1, 1, 1344) and anchors.shape == (1, 39, 3) and z.requires_grad
    ref = g['z_before'][3]
    close = np.abs(z.detach().numpy().reshape(-1) - ref) < 2e-4
    # Adam's first steps are sign-like: elements whose gradient is at round-off level may differ
    assert close.mean() > 0.98, close.mean()


def _joint_setup():
    from conftest import make_deformation
    g = load_golden('fit_joint.npz')
    lambdas = {'surface': 2.0, 'reg_expr': 0.01, 'reg_global': 0.25, 'reg_unobserved': 10, 'reg_loc': 0.05,
               'symm_dist': 5.0}
    schedule = {'lr': {200: 2, 400: 2, 600: 2, 800: 2}, 'symm_dist': {200: 10, 500: 9999},
                'reg_glob': {200: 3, 600: 10}, 'reg_loc': {500: 3, 600: 10}, 'reg_expr': {600: 10}}
    return g, lambdas, schedule, make_deformation


def test_joint_fitter_mirror_follows_reference():
    """2 iterations of inference_iterative_root_finding_joint (Broyden correspondences + implicit differentiation)
    through the mirrored modules on CPU, against the latents the reference produced."""
    from nphm_b200.models.fitting import inference_iterative_root_finding_joint
    g, lambdas, schedule, make_deformation = _joint_setup()
    dec = make_ensemble(0).train()
    dfn = make_deformation()
    np.random.seed(0)
    torch.manual_seed(0)
    z_ex, z_id, anchors = inference_iterative_root_finding_joint(dec, dfn, [torch.from_numpy(o) for o in g['obs']], lambdas,
                                                                 n_steps=200, schedule_cfg=schedule, step_scale=0.01)
    assert z_ex.shape == (3, 1, 200) and z_id.shape == (1, 1, 1344) and anchors.shape == (1, 39, 3)
    close_id = np.abs(z_id.detach().numpy().reshape(-1) - g['z_id_before'][2]) < 2e-4
    close_ex = np.abs(z_ex.detach().numpy().reshape(3, 200) - g['z_ex_before'][2]) < 2e-4
    assert close_id.mean() > 0.97 and close_ex.mean() > 0.97, (close_id.mean(), close_ex.mean())


def test_forward_mode_jacobian_equals_reverse_mode():
    """`jac` fast path (forward-mode through the DeformationNetwork) vs three reverse-mode passes (the reference's way)."""
    import torch
    from conftest import make_deformation
    from nphm_b200.models import diff_operators as D
    dfn = make_deformation()
    torch.manual_seed(1)
    xc = torch.randn(2, 40, 3) * 0.2
    cond = torch.randn(2, 1, 1344 + 200).repeat(1, 40, 1) * 0.1
    anchors = torch.randn(2, 40, 39, 3) * 0.1
    J_fast = D.jac(dfn, xc.clone(), cond, anchors)
    saved = dfn.offset_jacobian
    dfn.offset_jacobian = lambda *a, **k: None            # force the reverse-mode path
    try:
        J_ref = D.jac(dfn, xc.clone(), cond, anchors)
    finally:
        dfn.offset_jacobian = saved
    assert J_fast.shape == J_ref.shape == (2, 40, 3, 3)
    assert float((J_fast - J_ref).abs().max()) < 2e-6
    dfn.train()
    try:
        assert dfn.offset_jacobian(xc, cond, anchors) is None     # training-mode noise: reverse mode only
    finally:
        dfn.eval()


def test_search_mirror_follows_reference_single_and_multi_start():
    """The rewritten `search` / `broyden` (dense, flag-based) on CPU against the reference's own runs (search.npz):
    single start on 2 x 300 points and the multi-start default (5 starts, same torch seed) on 60 points."""
    import torch
    from test_oracle import _search_setup
    from nphm_b200.models.iterative_root_finding import search
    g, dfn, obs, cond, anchors = _search_setup()
    n = obs.shape[1]
    xc, res = search(obs, cond.repeat(1, n, 1), dfn, anchors, multi_corresp=False)
    valid = res['valid_ids'].numpy()
    both = valid & g['valid']
    assert xc.shape == (2, n, 3) and valid.shape == (2, n)
    assert (valid == g['valid']).mean() >= 0.97
    assert np.abs(xc.detach().numpy()[both] - g['xc'][both]).max() < 5e-5
    m = int(g['multi_n'])
    torch.manual_seed(int(g['multi_seed']))
    xm, rm = search(obs[:1, :m], cond[:1].repeat(1, m, 1), dfn, anchors[:1, :m], multi_corresp=True)
    vm = rm['valid_ids'].numpy()
    assert xm.shape == (1, m, 5, 3) and vm.shape == (1, m, 5)
    assert (vm == g['valid_multi']).mean() >= 0.97
    bothm = vm & g['valid_multi']
    assert np.abs(xm.detach().numpy()[bothm] - g['xc_multi'][bothm]).max() < 5e-5


def test_surface_loss_gradients_of_the_composite_match_reference_autograd():
    """Pins the chain  reference autograd -> composite module (here, CPU) -> nphm_fit_surface_grad (GPU test
    test_surface_loss_gradients_wrt_latent_and_points compares the kernel with this composite):  loss, d loss / d points
    and d loss / d latent of the joint fitter's surface term against surface_grad.npz (unmodified reference modules)."""
    import torch
    from conftest import load_golden, make_ensemble
    g = load_golden('surface_grad.npz')
    dec = make_ensemble(0).train()
    xc, valid = torch.from_numpy(g['points']), torch.from_numpy(g['valid'])
    for clamp in g['clamps']:
        tag = '%g' % clamp
        xa = xc.clone().requires_grad_(True)
        za = torch.from_numpy(g['latent']).reshape(1, 1, -1).clone().requires_grad_(True)
        sdf, _ = dec(xa, za.repeat(2, 1, 1), None)
        l = sdf[valid, :].abs()
        loss = l[l < float(clamp)].mean()
        loss.backward()
        assert int((l < float(clamp)).sum()) == int(g['kept_' + tag])
        assert abs(loss.item() - float(g['loss_' + tag])) < 1e-6
        gx, gz = xa.grad.numpy(), za.grad.reshape(-1).numpy()
        assert np.abs(gx - g['grad_points_' + tag]).max() < 2e-5 * np.abs(g['grad_points_' + tag]).max()
        assert np.abs(gz - g['grad_latent_' + tag]).max() < 2e-5 * np.abs(g['grad_latent_' + tag]).max()

"""Drop-in mirror of the reference module ``NPHM.models.diff_operators`` (the two helpers the hot path uses).

  * ``jac``       src/NPHM/models/diff_operators.py:26-54  - 3x3 Jacobian of x + F_ex(x) w.r.t. x
  * ``gradient``  src/NPHM/models/diff_operators.py:69-79  - spatial gradient of a scalar field (keeps the graph)

Like the reference, both set ``requires_grad_(True)`` on their input as a side effect.  For the drop-in
``DeformationNetwork`` ``jac`` evaluates the same Jacobian by forward-mode differentiation (no graph is returned; the
reference's callers - ``search`` and the joint fitter - detach it anyway).
"""
import torch


def jac(decoder_expr, xc, cond, anchors):
    """Returns d(xc + F_ex(xc)) / d xc as ``B x N x 3 x 3`` (row i = gradient of output i)."""
    xc.requires_grad_(True)
    fast = getattr(decoder_expr, 'offset_jacobian', None)
    if fast is not None:
        # forward-mode fast path of the drop-in DeformationNetwork (its callers only use the value of the Jacobian)
        with torch.no_grad():
            J = fast(xc.detach(), cond.detach(), anchors.detach() if anchors is not None else None)
        if J is not None:
            return J + torch.eye(3, device=J.device, dtype=J.dtype)
    xd, _ = decoder_expr(xc, cond, anchors)
    xd = xc + xd
    rows = []
    for i in range(xd.shape[-1]):
        seed = torch.zeros_like(xd)
        seed[..., i] = 1
        rows.append(torch.autograd.grad(outputs=xd, inputs=xc, grad_outputs=seed, create_graph=False,
                                        retain_graph=True, only_inputs=True)[0])
    return torch.stack(rows, dim=-2)


def gradient(outputs, inputs):
    """d outputs / d inputs summed over the output channel, last three input channels, graph kept."""
    seed = torch.ones_like(outputs)
    g = torch.autograd.grad(outputs=outputs, inputs=inputs, grad_outputs=seed, create_graph=True,
                            retain_graph=True, only_inputs=True, allow_unused=True)[0]
    return g[:, :, -3:]

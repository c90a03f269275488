"""Shared bounds of the parity tests (north star: gradients within 1e-4 relative, fp32)."""
import torch

RTOL = 1e-4


def rel_errors(a, b):
    """(relative L2 error, max-norm error) of a against b, in float64."""
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    l2 = float((a - b).norm() / b.norm().clamp_min(1e-300))
    mx = float((a - b).abs().max() / b.abs().max().clamp_min(1e-300))
    return l2, mx


def close_grad(a, b, what, rtol=RTOL):
    """Gradient tensors: max |a-b| <= rtol max|b| and ||a-b||_2 <= rtol ||b||_2."""
    assert a.shape == b.shape, f"{what}: shape {tuple(a.shape)} vs {tuple(b.shape)}"
    l2, mx = rel_errors(a, b)
    assert l2 <= rtol and mx <= rtol, f"{what}: relative L2 error {l2:.3e}, max-norm error {mx:.3e} > {rtol}"
    return l2, mx


def adam_first_step_violations(p_mine, p_before, g_ref, lr, rtol=RTOL, eps=1e-8, masked=False, slack=2e-6):
    """ELEMENT-WISE bound on the first MaskedAdam step (model/cuda/adam_upd_kernel.cu:8-40, betas 0.9 / 0.99).

    With zero moments the update of an element is u(g) = -lr * g / (|g| + eps'), eps' = eps / sqrt(1 - beta2) once
    step_size = lr sqrt(1 - beta2) / (1 - beta1) is folded in: a function of that element's gradient only, monotone in g, saturating at -+lr.  A gradient
    tensor within `rtol` of the reference's in max norm means |g - g_ref| <= delta = rtol * max|g_ref| per element,
    so our update must lie in [u(g_ref + delta), u(g_ref - delta)] (+- slack * lr for fp32 rounding of p itself).
    Where |g_ref| >> delta this pins the update to ~1e-4 relative; where the gradient is rounding noise it allows
    what the 1e-4 gradient bar allows and no more.  Masked Adam: elements with g_ref == 0 must not move at all
    unless our gradient may legitimately be non-zero there (never, for bit-identical sample sets).
    Returns the number of elements outside their interval."""
    p_mine, p_before, g = p_mine.detach().double(), p_before.detach().double(), g_ref.detach().double()
    delta = rtol * float(g.abs().max())
    e = eps / (1 - 0.99) ** 0.5

    def u(x):
        return -lr * x / (x.abs() + e)
    lo, hi = u(g + delta), u(g - delta)              # u is decreasing in g
    d = p_mine - p_before
    tol = slack * lr + 1e-7 * p_before.abs()         # one fp32 ulp of the parameter
    bad = (d < lo - tol) | (d > hi + tol)
    if masked:
        untouched = g == 0
        bad = torch.where(untouched, d != 0, bad)
    return int(bad.sum())

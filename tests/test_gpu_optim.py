"""Fused clip + Adam and the gated modality scatter against the torch ops the reference's task loops
and MRGCN._compute_modality_embeddings run (node_classification.py:190-193, tasks/utils.py:8-45, mrgcn.py:295-301)."""
import pytest
import torch

from parity import check

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _adam_f64(params, grads_per_step, lr, betas, eps, wd, max_norm):
    """float64 evaluation of clip_grad_norm_ + Adam (the adjudicator)."""
    p = [x.double().clone() for x in params]
    m = [torch.zeros_like(x) for x in p]
    v = [torch.zeros_like(x) for x in p]
    for t, grads in enumerate(grads_per_step, 1):
        g = [x.double() for x in grads]
        total = torch.sqrt(sum((x * x).sum() for x in g))
        coef = min(1.0, max_norm / (float(total) + 1e-6))
        for i in range(len(p)):
            gi = g[i] * coef + wd * p[i]
            m[i] = betas[0] * m[i] + (1 - betas[0]) * gi
            v[i] = betas[1] * v[i] + (1 - betas[1]) * gi * gi
            p[i] = p[i] - lr / (1 - betas[0] ** t) * m[i] / (v[i].sqrt() / (1 - betas[1] ** t) ** 0.5 + eps)
    return p


@pytest.mark.parametrize("wd", [0.0, 0.01])
def test_fused_clip_adam_matches_torch(wd):
    from mrgcn_b200.optim import FusedClipAdam
    torch.manual_seed(0)
    shapes = [(40 * 5000, 10), (267, 40), (40, 151, 10), (11,), (3, 7)]
    init = [torch.randn(s) * 0.1 for s in shapes]
    grads = [[torch.randn(s) * (3.0 if k == 0 else 0.01) for s in shapes] for k in range(3)]      # step 1 is clipped hard
    lr, betas, eps, max_norm = 0.01, (0.9, 0.999), 1e-8, 1.0
    ref_p = [torch.nn.Parameter(x.clone().to(DEV)) for x in init]
    ref_opt = torch.optim.Adam([{"params": ref_p[:2]}, {"params": ref_p[2:], "lr": lr}], lr=lr, betas=betas, eps=eps, weight_decay=wd)
    our_p = [torch.nn.Parameter(x.clone().to(DEV)) for x in init]
    our_opt = FusedClipAdam([{"params": our_p[:2]}, {"params": our_p[2:], "lr": lr}], lr=lr, betas=betas, eps=eps, weight_decay=wd,
                            max_norm=max_norm)
    for g in grads:
        for p, q, gi in zip(ref_p, our_p, g):
            p.grad = gi.clone().to(DEV)
            q.grad = gi.clone().to(DEV)
        torch.nn.utils.clip_grad_norm_(ref_p, max_norm)
        ref_opt.step()
        our_opt.step()
    tru = _adam_f64(init, grads, lr, betas, eps, wd, max_norm)
    for k, (p, q, t) in enumerate(zip(ref_p, our_p, tru)):
        check(q, p, t, "fused clip+adam wd=%g tensor %d" % (wd, k))
    # the squared-norm kernel against torch
    tot = our_opt.total_sqnorm()
    want = sum(float(q.grad.double().pow(2).sum()) for q in our_p)
    assert abs(float(tot) - want) <= 1e-12 * want


def test_gated_scatter_matches_index_assignment():
    from mrgcn_b200.optim import gated_scatter
    torch.manual_seed(1)
    N, D, d, m = 5000, 23, 7, 1234
    rows = torch.sort(torch.randperm(N)[:m]).values
    src = torch.randn(m, d, requires_grad=True)
    gate = torch.tensor(0.1, requires_grad=True)
    G = torch.randn(N, D)
    # reference ops (mrgcn.py:295-301)
    Xr = torch.zeros(N, D)
    mask = torch.zeros(N, dtype=torch.bool)
    mask[rows] = True
    Xr[mask, 5:5 + d] = torch.mul(src, gate)
    (Xr * G).sum().backward()
    src_c = src.detach().to(DEV).requires_grad_(True)
    gate_c = gate.detach().to(DEV).requires_grad_(True)
    X = gated_scatter(torch.zeros(N, D, device=DEV), src_c, rows, gate_c, 5)
    check(X, Xr, None, "gated scatter values")
    (X * G.to(DEV)).sum().backward()
    check(src_c.grad, src.grad, None, "gated scatter grad src")
    check(gate_c.grad.reshape(1), gate.grad.reshape(1), (src.double() * G[rows, 5:5 + d].double()).sum().reshape(1), "gated scatter grad gate")

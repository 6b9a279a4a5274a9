"""Fused gradient clipping + Adam — what the reference's task loops do with
`nn.utils.clip_grad_norm_(model.parameters(), 1.0)` followed by `optimizer.step()` of a stock `torch.optim.Adam`
(/root/reference/mrgcn/tasks/node_classification.py:190-193, link_prediction.py:324-326; optimizer groups built by
mrgcn/tasks/utils.py:8-45), in two passes over the gradients instead of ~14 passes over the 2.67 GB identity table of AM.

    opt = FusedClipAdam(param_groups, max_norm=1.0)        # same param_groups / lr / betas / eps / weight_decay as optim.Adam
    loss.backward(); opt.step()                            # clipping happens inside step(): drop the clip_grad_norm_ call

The arithmetic is torch.optim.Adam's (no amsgrad) on gradients scaled by clip_grad_norm_'s coefficient.  With a node
partition (`sharded=[weight_I shard]`) the squared norm of the sharded tensors is summed over the ranks with one 8-byte
all-reduce; replicated tensors (already identical on every rank after the gradient all-reduce) count once.
CUDA tensors go through the C ABI (csrc/optim.cu); tensors on the CPU (the reference keeps literal encoders there) take
the same formulas through torch ops.
"""
from __future__ import annotations

import math

import torch
import torch.distributed as dist

from . import _native as nv


class FusedClipAdam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, max_norm=None, sharded=(), group=None):
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        self.max_norm = max_norm
        self._sharded = {id(p) for p in sharded}
        self._group = group
        self._ws = {}

    def _scratch(self, dev):
        if dev not in self._ws:
            self._ws[dev] = (torch.empty(int(nv.lib().mrgcn_sqnorm_ws_elems()), dtype=torch.float64, device=dev),
                             torch.zeros(2, dtype=torch.float64, device=dev))
        return self._ws[dev]

    @torch.no_grad()
    def total_sqnorm(self):
        """Sum of squared gradients over every parameter of every group (the quantity clip_grad_norm_ takes the root of),
        as a float64 tensor on the device of the first CUDA parameter (or the CPU)."""
        ps = [p for g in self.param_groups for p in g["params"] if p.grad is not None]
        cuda = [p for p in ps if p.is_cuda]
        dev = cuda[0].device if cuda else torch.device("cpu")
        total = torch.zeros(2, dtype=torch.float64, device=dev)      # [replicated, sharded]
        for p in ps:
            slot = 1 if id(p) in self._sharded else 0
            g = p.grad
            if g.is_cuda and g.dtype == torch.float32 and g.is_contiguous() and g.data_ptr() % 16 == 0:
                ws, acc = self._scratch(g.device)
                with torch.cuda.device(g.device):
                    nv.check(nv.lib().mrgcn_grad_sqnorm(g.data_ptr(), g.numel(), ws.data_ptr(), acc.data_ptr(), 0,
                                                        nv.stream_ptr()), "grad_sqnorm")
                total[slot] += acc[0].to(dev)
            else:
                total[slot] += g.double().pow(2).sum().to(dev)
        if self._sharded and dist.is_available() and dist.is_initialized() and dist.get_world_size(self._group) > 1:
            dist.all_reduce(total[1:], group=self._group)
        return total.sum().reshape(1)

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        total = self.total_sqnorm() if self.max_norm is not None else None
        for group in self.param_groups:
            b1, b2 = group["betas"]
            for p in group["params"]:
                if p.grad is None:
                    continue
                st = self.state[p]
                if not st:
                    st["step"] = 0
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                st["step"] += 1
                g = p.grad
                fused = (p.is_cuda and p.dtype == torch.float32 and p.is_contiguous() and g.is_contiguous()
                         and all(t.data_ptr() % 16 == 0 for t in (p, g, st["exp_avg"], st["exp_avg_sq"])))
                if fused:
                    tot = total.to(p.device) if total is not None else None
                    with torch.cuda.device(p.device):
                        nv.check(nv.lib().mrgcn_adam_clip(p.data_ptr(), g.data_ptr(), st["exp_avg"].data_ptr(),
                                                          st["exp_avg_sq"].data_ptr(), p.numel(), nv.ptr(tot),
                                                          float(self.max_norm or 0.0), group["lr"], b1, b2, group["eps"],
                                                          group["weight_decay"], st["step"], nv.stream_ptr()), "adam_clip")
                    continue
                # same arithmetic through torch ops (CPU tensors, odd layouts)
                coef = 1.0
                if total is not None:
                    coef = min(1.0, float(self.max_norm) / (math.sqrt(float(total)) + 1e-6))
                gg = g * coef
                if group["weight_decay"]:
                    gg = gg.add(p, alpha=group["weight_decay"])
                st["exp_avg"].lerp_(gg, 1 - b1)
                st["exp_avg_sq"].mul_(b2).addcmul_(gg, gg, value=1 - b2)
                bc1, bc2 = 1 - b1 ** st["step"], 1 - b2 ** st["step"]
                denom = (st["exp_avg_sq"].sqrt() / math.sqrt(bc2)).add_(group["eps"])
                p.addcdiv_(st["exp_avg"], denom, value=-group["lr"] / bc1)
        return loss


class _GatedScatter(torch.autograd.Function):
    """X[row_idx, col0:col0+d] = gate * src, in place (mrgcn/models/mrgcn.py:295-301: `out = torch.mul(out, gate)` then the
    masked assignment into the zero feature matrix); one kernel forward, one backward."""

    @staticmethod
    def forward(ctx, X, src, row_idx, gate, col0):
        src = src.contiguous().float()
        gate1 = gate.reshape(1).to(src.device).float()
        with torch.cuda.device(X.device):
            nv.check(nv.lib().mrgcn_scatter_rows(src.data_ptr(), row_idx.data_ptr(), gate1.data_ptr(), X.data_ptr(), src.shape[0],
                                                 src.shape[1], X.stride(0), int(col0), nv.stream_ptr()), "scatter_rows")
        ctx.mark_dirty(X)
        ctx.save_for_backward(src, row_idx, gate1)
        ctx.col0, ctx.gate_shape, ctx.gate_dev = int(col0), gate.shape, gate.device
        return X

    @staticmethod
    def backward(ctx, gX):
        src, row_idx, gate1 = ctx.saved_tensors
        gX = gX.contiguous()
        g_src = torch.empty_like(src)
        g_gate = torch.empty(1, dtype=torch.float32, device=src.device)
        ws = torch.empty(int(nv.lib().mrgcn_sqnorm_ws_elems()), dtype=torch.float64, device=src.device)
        with torch.cuda.device(src.device):
            nv.check(nv.lib().mrgcn_scatter_rows_bwd(gX.data_ptr(), row_idx.data_ptr(), gate1.data_ptr(), src.data_ptr(),
                                                     g_src.data_ptr(), g_gate.data_ptr(), ws.data_ptr(), src.shape[0], src.shape[1],
                                                     gX.stride(0), ctx.col0, nv.stream_ptr()), "scatter_rows_bwd")
        # the written block of X is overwritten, not accumulated: no gradient flows to its previous (zero) contents there
        gXin = gX.clone()
        gXin[row_idx, ctx.col0:ctx.col0 + src.shape[1]] = 0
        return gXin, g_src, None, g_gate.reshape(ctx.gate_shape).to(ctx.gate_dev), None


def gated_scatter(X, src, row_idx, gate, col0):
    """In-place X[row_idx, col0:col0+d] = gate * src on a CUDA feature matrix X (returns X, autograd-connected)."""
    nv.require_cuda(X, "X")
    return _GatedScatter.apply(X, src.to(X.device), row_idx.to(X.device).long().contiguous(), gate, col0)

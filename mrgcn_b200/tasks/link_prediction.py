"""Link-prediction hot path: DistMult scorer, in-batch negative sampling and raw/filtered ranking —
drop-ins for the module-level functions of /root/reference/mrgcn/tasks/link_prediction.py
(`score_distmult_bc` :645-665, negative sampling :244-268, `compute_ranks_fast` :593-643,
`filter_scores_` :557-573, `truedicts` :576-591).  CUDA only."""
from __future__ import annotations

import numpy as np
import torch

from .. import _native as nv


class _DistMultFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, s, p, o, E, Rel):
        for t, n in ((s, "s"), (p, "p"), (o, "o"), (E, "node_embeddings"), (Rel, "edge_embeddings")):
            nv.require_cuda(t, n)
        E, Rel = E.contiguous(), Rel.contiguous()
        n, h = s.numel(), E.shape[1]
        score = torch.empty(n, dtype=torch.float32, device=E.device)
        with torch.cuda.device(E.device):
            nv.check(nv.lib().mrgcn_distmult_fwd(nv.ptr(s), nv.ptr(p), nv.ptr(o), n, nv.ptr(E), nv.ptr(Rel), h,
                                                 nv.ptr(score), nv.stream_ptr()), "distmult_fwd")
        ctx.save_for_backward(s, p, o, E, Rel)
        return score

    @staticmethod
    def backward(ctx, g):
        s, p, o, E, Rel = ctx.saved_tensors
        n, h = s.numel(), E.shape[1]
        g = g.contiguous().float()
        gE = torch.empty_like(E) if ctx.needs_input_grad[3] else None
        gRel = torch.empty_like(Rel) if ctx.needs_input_grad[4] else None
        ws = torch.empty(int(nv.lib().mrgcn_distmult_bwd_ws_elems(n)), dtype=torch.int32, device=E.device)
        with torch.cuda.device(E.device):
            nv.check(nv.lib().mrgcn_distmult_bwd(nv.ptr(s), nv.ptr(p), nv.ptr(o), n, nv.ptr(g), nv.ptr(E), nv.ptr(Rel),
                                                 E.shape[0], Rel.shape[0], h, nv.ptr(gE), nv.ptr(gRel), nv.ptr(ws),
                                                 nv.stream_ptr()), "distmult_bwd")
        return None, None, None, gE, gRel


def score_distmult_bc(data, node_embeddings, edge_embeddings):
    """score = sum_k E[s,k] * Rel[p,k] * E[o,k] for index tensors of any common (broadcastable) shape
    (link_prediction.py:645-665; the three matmul short-cuts of :652-663 compute the same numbers)."""
    si, pi, oi = data
    dev = node_embeddings.device
    si, pi, oi = torch.broadcast_tensors(torch.as_tensor(si).to(dev).long(), torch.as_tensor(pi).to(dev).long(),
                                         torch.as_tensor(oi).to(dev).long())
    shape = si.shape              # every path of the reference returns the broadcast index shape
    out = _DistMultFn.apply(si.reshape(-1).contiguous(), pi.reshape(-1).contiguous(), oi.reshape(-1).contiguous(),
                            node_embeddings.float(), edge_embeddings.to(dev).float())
    return out.view(shape)


def negative_samples(batch_data, rng=np.random):
    """link_prediction.py:244-268: corrupt floor(n/5) triples chosen without replacement, first half heads,
    second half tails, replacements drawn from the batch's own node set; labels 1 (true) / 0 (corrupted).
    Host NumPy RNG on purpose: same calls in the same order as the reference, so the same seed gives the
    same corrupted triples."""
    batch_data = np.asarray(batch_data)
    n = batch_data.shape[0]
    nodes = np.union1d(batch_data[:, 0], batch_data[:, 2])
    ncorrupt = n // 5
    pick = rng.choice(np.arange(n), ncorrupt, replace=False)
    nhead = ncorrupt // 2
    ntail = ncorrupt - nhead
    corrupted = np.empty((ncorrupt, 3), dtype=int)
    corrupted[:] = batch_data[pick]
    corrupted[:nhead, 0] = rng.choice(nodes, nhead)
    corrupted[-ntail:, 2] = rng.choice(nodes, ntail)
    Y = torch.ones(n + ncorrupt, dtype=torch.float32)
    Y[-ncorrupt:] = 0
    return corrupted, Y


def truedicts(facts):
    """link_prediction.py:576-591."""
    heads, tails = dict(), dict()
    for s, p, o in np.asarray(facts).tolist():
        heads.setdefault((p, o), []).append(s)
        tails.setdefault((s, p), []).append(o)
    return heads, tails


def filter_csr_device(facts, head):
    """The reference's filter dictionaries (`truedicts` + `filter_scores_`, link_prediction.py:557-591) as a CSR over the
    facts, built on the device by sort + segment: for fact f the list of all subjects s' with (s', p_f, o_f) in `facts`
    (head side) or all objects o' with (s_f, p_f, o') in `facts` (tail side).  The target itself stays in the list (the
    kernel skips it, as filter_scores_ does).  facts: int64 (F, 3) device tensor.  Returns int32 (ptr[F+1], idx)."""
    F = facts.shape[0]
    dev = facts.device
    if F == 0:
        z = torch.zeros(1, dtype=torch.int32, device=dev)
        return z, z
    a, b = (facts[:, 1], facts[:, 2]) if head else (facts[:, 0], facts[:, 1])
    member = facts[:, 0] if head else facts[:, 2]
    # group of a fact = its fixed pair (dense ids via unique); the group's member list is the set of distinct members
    # (the reference's lists keep duplicates of repeated facts, which set the same score to -inf twice: same result)
    key = a * (int(b.max()) + 1) + b
    _, inv = torch.unique(key, return_inverse=True)
    M = int(member.max()) + 1
    pairs = torch.unique(inv * M + member)                        # sorted: grouped by key, members ascending
    grp = pairs // M
    members_sorted = pairs - grp * M
    sizes = torch.bincount(grp, minlength=int(inv.max()) + 1)
    start = torch.cumsum(sizes, 0) - sizes
    cnt = sizes[inv]                                              # list length of every fact = size of its group
    ptr = torch.zeros(F + 1, dtype=torch.long, device=dev)
    torch.cumsum(cnt, 0, out=ptr[1:])
    total = int(ptr[-1])
    assert total < 2 ** 31, "filter lists exceed int32 indexing; rank the facts in several calls"
    fact_of = torch.repeat_interleave(torch.arange(F, device=dev), cnt)
    pos = torch.arange(total, device=dev) - ptr[fact_of] + start[inv[fact_of]]
    idx = members_sorted[pos]
    return ptr.to(torch.int32), idx.to(torch.int32).contiguous()


RANK_MAX_FACTS = 1 << 20      # facts per ranking call (the C entry point takes 65 535 * 32)


def compute_ranks_fast(data, node_embeddings, edge_embeddings, batch_size=16, filtered=True):
    """Ranks of every fact against all N candidate tails, then all candidate heads (loop order of
    link_prediction.py:602).  `batch_size` (the reference's mrr_batchsize chunking of the fact axis, :618-625) does not
    change any result and is not needed here: the scores live in registers only (csrc/rank.cu).  The filter lists are
    built over ALL facts of `data`, as the reference does (:597-600), on the device.  Returns int64 (2*facts,), 1-based."""
    E = node_embeddings.detach().float().contiguous()
    nv.require_cuda(E, "node_embeddings")
    dev = E.device
    Rel = edge_embeddings.detach().to(dev).float().contiguous()
    facts = torch.as_tensor(data).to(dev).long().contiguous()
    F, N, h = facts.shape[0], E.shape[0], E.shape[1]
    out = torch.empty(2 * F, dtype=torch.int64, device=dev)
    for k, head in enumerate((False, True)):
        fptr = fidx = None
        if filtered:
            fptr, fidx = filter_csr_device(facts, head)
        for f0 in range(0, F, RANK_MAX_FACTS):
            f1 = min(F, f0 + RANK_MAX_FACTS)
            ws = torch.empty(int(nv.lib().mrgcn_distmult_rank_ws_elems(f1 - f0)), dtype=torch.int32, device=dev)
            fp = None
            if filtered:
                fp = (fptr[f0:f1 + 1] - fptr[f0]).contiguous()
                fi = fidx[int(fptr[f0]):] if f0 else fidx
            with torch.cuda.device(dev):
                nv.check(nv.lib().mrgcn_distmult_rank(nv.ptr(facts[f0:f1]), f1 - f0, int(head), nv.ptr(E), nv.ptr(Rel), N, h,
                                                      nv.ptr(fp), nv.ptr(fi) if filtered else None, nv.ptr(ws),
                                                      nv.ptr(out[k * F + f0:]), nv.stream_ptr()), "distmult_rank")
    return out

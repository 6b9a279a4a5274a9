"""CPU oracle for the R-GCN / DistMult hot path.  TEST INFRASTRUCTURE ONLY.

This module restates, op for op, what wxwilcke/mrgcn computes on the path named in
BASELINE.json (`north_star`), using the same scipy / torch-CPU library calls the reference
makes, so its floating-point results are the reference's results.  Only `tests/`,
`__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference` legs of `bench.py` may
import it.  Nothing under `mrgcn_b200/` imports it; the product path has no CPU fallback.

Parity status: PINNED.  The reference holds no golden vectors for this path (SURVEY.md §4,
§8c), so the oracle is pinned against outputs of the reference itself: `tests/golden/
make_golden.py` imports the unmodified reference from /root/reference, runs it on seeded
inputs and stores inputs + outputs under `tests/golden/*.npz`; `tests/test_oracle_golden.py`
checks every function below against those files.

Every function cites the reference lines (relative to /root/reference) it follows.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp
import torch


# --------------------------------------------------------------------------------------
# adjacency construction  (mrgcn/encodings/graph_structure.py)
# --------------------------------------------------------------------------------------
def normalize_rows(adj):
    """Row-normalise one relation block: val = 1 / rowsum, empty rows stay empty.
    Follows graph_structure.py:162-169 (sum over axis 1, reciprocal, inf -> 0, diag @ adj)."""
    with np.errstate(divide="ignore"):
        deg = np.asarray(adj.sum(1)).ravel()
        inv = 1.0 / deg
        inv[np.isinf(inv)] = 0.0
    return sp.diags(inv).dot(adj).tocsr()


def stacked_adjacency(triples, num_nodes, num_props, include_inverse=True):
    """Integer triples (s, p, o) -> the N x (R*N) CSR the reference stores in its tarball.

    Relation block 2p holds A[s, o], block 2p+1 the inverse A[o, s] (graph_structure.py:78-106,
    blocks appended in sorted-property order), the last block is the identity
    (graph_structure.py:33-35) and everything is hstack'ed to CSR (graph_structure.py:38).
    Properties without triples still get (empty) blocks here because ids are dense.
    The float64 result is what `Tarball._store_csr` writes; `as_float32` below is the cast
    `Tarball._read_csr` (tarball.py:151-157) applies on load.
    """
    triples = np.asarray(triples)
    shape = (num_nodes, num_nodes)
    blocks = []
    for p in range(num_props):
        sel = triples[triples[:, 1] == p]
        row = sel[:, 0].astype(np.int32)
        col = sel[:, 2].astype(np.int32)
        ones = np.ones(len(row), dtype=np.int8)
        blocks.append(normalize_rows(sp.csr_matrix((ones, (row, col)), shape=shape, dtype=np.int8)))
        if include_inverse:
            blocks.append(normalize_rows(sp.csr_matrix((ones, (col, row)), shape=shape, dtype=np.int8)))
    blocks.append(normalize_rows(sp.identity(num_nodes).tocsr()))
    return sp.hstack(blocks, format="csr")


def as_float32(A):
    """tarball.py:151-157 — the stacked adjacency is re-read as float32."""
    return sp.csr_matrix((A.data.astype(np.float32), A.indices, A.indptr), shape=A.shape)


def csr_to_coo(A, dtype=torch.int8):
    """scipy CSR -> torch sparse COO exactly as data/utils.py:165-170 does (indices from
    `.nonzero()`, values from `.data`, then a dtype cast; FullBatch passes int8, batch.py:148)."""
    idx = np.array(A.nonzero())
    return torch.sparse_coo_tensor(torch.LongTensor(idx), torch.Tensor(A.data), A.shape, dtype=dtype)


def node_column_index(idx, num_nodes, num_relations):
    """batch.py:245-250 — column ids r*N + i for every relation r and node i in `idx`."""
    idx = torch.as_tensor(idx, dtype=torch.int64)
    rel = torch.arange(num_relations, dtype=torch.int64).view(-1, 1)
    return (rel * num_nodes + idx.view(1, -1)).reshape(-1)


def slice_columns(A, col_idx):
    """batch.py:252-263 — keep the entries whose column is in `col_idx`, renumber columns by
    position in `col_idx`, and reset every kept value to 1.0 (float32)."""
    ind = A._indices()
    keep = torch.isin(ind[1], col_idx)
    row, col = ind[0][keep], ind[1][keep]
    lut = torch.full((int(A.shape[1]),), -1, dtype=torch.int64)
    lut[col_idx] = torch.arange(len(col_idx), dtype=torch.int64)
    return torch.sparse_coo_tensor(torch.vstack([row, lut[col]]),
                                   torch.ones(len(col), dtype=torch.float32),
                                   size=[A.shape[0], len(col_idx)])


# --------------------------------------------------------------------------------------
# parameters  (mrgcn/layers/graph.py:9-60,104-116 ; mrgcn/models/rgcn.py:12-61,130-132)
# --------------------------------------------------------------------------------------
def init_layer_params(indim, outdim, num_relations, num_nodes, num_bases=-1, bias=False,
                      input_layer=False, featureless=False):
    """Same tensors, shapes, registration order and initialisers as GraphConvolution.__init__ /
    reset_parameters: comps first, then weight_I, weight_F, b; Xavier-uniform on all but b."""
    p = {}
    S = num_relations
    if num_bases > 0:
        S = num_bases
        if input_layer:
            p["weight_I_comp"] = torch.empty(num_relations, num_bases)
        if not featureless:
            p["weight_F_comp"] = torch.empty(num_relations, num_bases)
    if input_layer:
        p["weight_I"] = torch.empty(S * num_nodes, outdim)
    if not featureless:
        p["weight_F"] = torch.empty(S, indim, outdim)
    if bias:
        p["b"] = torch.empty(outdim)
    for name, t in p.items():
        if name == "b":
            continue
        torch.nn.init.xavier_uniform_(t)
    if bias:
        torch.nn.init.zeros_(p["b"])
    return p


def init_rgcn_params(modules, num_relations, num_nodes, num_bases, featureless, bias,
                     link_prediction):
    """rgcn.py:12-61 — layer_0 is the input layer, the rest are hidden; `relations` last."""
    layers = []
    for k, (indim, outdim, _ltype, _act) in enumerate(modules):
        layers.append(init_layer_params(indim, outdim, num_relations, num_nodes, num_bases, bias,
                                        input_layer=(k == 0),
                                        featureless=(featureless if k == 0 else False)))
    relations = None
    if link_prediction:
        relations = torch.empty(num_relations, modules[-1][1])
        torch.nn.init.xavier_uniform_(relations)
    return layers, relations


# --------------------------------------------------------------------------------------
# the layer  (mrgcn/layers/graph.py:62-102)
# --------------------------------------------------------------------------------------
def graphconv_forward(p, X, A, *, num_nodes, num_relations, num_bases, input_layer, featureless,
                      A_idx=None, dtype=torch.float32):
    """One R-GCN layer, same op sequence as GraphConvolution.forward.
    dtype: torch.float32 is the reference (`A.float()`, graph.py:75,95); torch.float64 with double parameters / X
    evaluates the same op sequence in double precision — the "truth" the parity tests adjudicate fp32 rounding
    differences against (tests/parity.py)."""
    outdim = (p["weight_I"] if "weight_I" in p else p["weight_F"]).shape[-1]
    ident = 0.0
    if input_layer:
        W_I = p["weight_I"]
        if num_bases > 0:                                                   # graph.py:69-72
            W_I = torch.einsum("rb,bij->rij", p["weight_I_comp"],
                               W_I.view(num_bases, num_nodes, outdim))
            W_I = W_I.view(num_relations * num_nodes, outdim)
        ident = torch.mm(A.to(dtype), W_I)                                    # graph.py:75
        if featureless:                                                     # graph.py:77-81
            return ident + p["b"] if "b" in p else ident
    W_F = p["weight_F"]
    if num_bases > 0:                                                       # graph.py:83-85
        W_F = torch.einsum("rb,bij->rij", p["weight_F_comp"], W_F)
    n = num_nodes
    if A_idx is not None:                                                   # graph.py:88-91
        n = X.shape[0]
        A = slice_columns(A, A_idx)
    proj = torch.einsum("ij,bjk->bik", X, W_F).reshape(num_relations * n, outdim)   # :93-94
    feat = torch.mm(A.to(dtype), proj)                                        # graph.py:95
    out = ident + feat if input_layer else feat                             # graph.py:97
    if "b" in p:
        out = out + p["b"]                                                  # graph.py:99-100
    return out


def rgcn_forward(layers, activations, X, A, *, num_nodes, num_relations, num_bases, featureless,
                 row_masks=None, dtype=torch.float32):
    """rgcn.py:69-89, full batch.  `row_masks[k]` (shape (N,), already scaled by 1/(1-p)) stands
    in for the dropout-on-ones vector of rgcn.py:82-84, whose RNG stream is not reproducible."""
    for k, (p, act) in enumerate(zip(layers, activations)):
        X = graphconv_forward(p, X, A, num_nodes=num_nodes, num_relations=num_relations,
                              num_bases=num_bases, input_layer=(k == 0),
                              featureless=(featureless if k == 0 else False), dtype=dtype)
        if row_masks is not None and row_masks[k] is not None:
            X = torch.mul(X.T, row_masks[k]).T
        if act == "relu":
            X = torch.relu(X)
    return X


# --------------------------------------------------------------------------------------
# MRGCN: gated scatter of literal-encoder outputs  (mrgcn/models/mrgcn.py:189-214,250-305)
# --------------------------------------------------------------------------------------
def mlp_forward(weights, x):
    """mrgcn/models/perceptron.py:6-46 with p_dropout = 0: Linear -> (Dropout) -> ReLU per layer.
    weights: [(W, b), ...] in layer order."""
    for W, b in weights:
        x = torch.relu(torch.nn.functional.linear(x, W, b))
    return x


def modality_features(num_nodes, sets, gate_weights, dtype=torch.float32):
    """mrgcn.py:250-305 for a full batch: zeros (N, sum dim); per encoding set (in order) the encoder output times its
    gate weight is written into the rows listed in node_idx.  sets: [(mlp_weights, encodings, node_idx), ...]."""
    cols = []
    for i, (weights, enc, node_idx) in enumerate(sets):
        out = mlp_forward(weights, enc.to(dtype)) * gate_weights[i]
        block = torch.zeros((num_nodes, out.shape[1]), dtype=dtype)
        block = block.index_put((torch.as_tensor(node_idx),), out)
        cols.append(block)
    return torch.cat(cols, dim=1)


# --------------------------------------------------------------------------------------
# link prediction  (mrgcn/tasks/link_prediction.py)
# --------------------------------------------------------------------------------------
def distmult_score(idx, node_emb, rel_emb):
    """link_prediction.py:645-665, including the three broadcast short-cuts."""
    si, pi, oi = idx
    s, p, o = node_emb[si, :], rel_emb[pi, :], node_emb[oi, :]
    if s.dim() == p.dim() == o.dim():
        if pi.size(-1) == 1 and oi.size(-1) == 1:
            return torch.matmul(s, (p * o).transpose(-1, -2)).squeeze(-1)
        if si.size(-1) == 1 and oi.size(-1) == 1:
            return torch.matmul(p, (s * o).transpose(-1, -2)).squeeze(-1)
        if si.size(-1) == 1 and pi.size(-1) == 1:
            return torch.matmul(o, (s * p).transpose(-1, -2)).squeeze(-1)
    return torch.sum(s * p * o, dim=-1)


def negative_samples(batch_data, rng):
    """link_prediction.py:244-268 with an explicit numpy RandomState-like `rng`
    (the reference uses the global np.random stream; same calls, same order)."""
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
    labels = torch.ones(n + ncorrupt, dtype=torch.float32)
    labels[-ncorrupt:] = 0
    return corrupted, labels


def true_dicts(facts):
    """link_prediction.py:576-591."""
    heads, tails = {}, {}
    for s, p, o in np.asarray(facts).tolist():
        heads.setdefault((p, o), []).append(s)
        tails.setdefault((s, p), []).append(o)
    return heads, tails


def compute_ranks(data, node_emb, rel_emb, mrr_batchsize, filtered=True):
    """link_prediction.py:593-643 incl. filter_scores_ (:557-573): tail side first, then head;
    chunking over the FACT axis with bounds derived from num_nodes (the reference's quirk)."""
    data = torch.as_tensor(data)
    heads, tails = true_dicts(data) if filtered else (None, None)
    nf, nn_ = data.shape[0], node_emb.shape[0]
    out = torch.empty(nf * 2, dtype=torch.int64)
    off = 0
    for head in (False, True):
        bases = data[:, 1:] if head else data[:, :2]
        targets = data[:, 0] if head else data[:, 2]
        bexp = bases.view(nf, 1, 2).expand(nf, nn_, 2)
        ar = torch.arange(nn_).view(1, nn_, 1).expand(nf, nn_, 1)
        cand = torch.cat([ar, bexp] if head else [bexp, ar], dim=2)
        scores = torch.zeros(cand.shape[:2])
        for b0 in range(0, nn_, mrr_batchsize):
            sl = slice(b0, min(b0 + mrr_batchsize, nn_))
            scores[sl] = distmult_score((cand[sl, :, 0], cand[sl, :, 1], cand[sl, :, 2]),
                                        node_emb, rel_emb)
        if filtered:
            hit = []
            for i, (s, p, o) in enumerate(data.tolist()):
                if head:
                    hit.extend((i, x) for x in heads[p, o] if x != s)
                else:
                    hit.extend((i, x) for x in tails[s, p] if x != o)
            if hit:
                hit = torch.tensor(hit)
                scores[hit[:, 0], hit[:, 1]] = float("-inf")
        true = scores[torch.arange(nf), targets.long()].view(nf, 1)
        ranks = torch.sum(scores > true, dim=1, dtype=torch.int64)
        ties = torch.sum(scores == true, dim=1, dtype=torch.int64)
        out[off:off + nf] = ranks + torch.round((ties - 1) / 2).long()
        off += nf
    return out + 1


# --------------------------------------------------------------------------------------
# losses  (node_classification.py:439-444 ; link_prediction.py:550-554)
# --------------------------------------------------------------------------------------
def nc_loss(logits, labelled_idx, targets):
    return torch.nn.functional.cross_entropy(logits[labelled_idx], targets)


def lp_loss(scores, labels):
    return torch.nn.functional.binary_cross_entropy_with_logits(scores, labels)

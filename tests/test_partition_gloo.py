"""Node-partitioned R-GCN (mrgcn_b200/partition.py) on 2 CPU ranks over gloo: the partition logic and the collective
autograd functions are the product's; the per-rank layer arithmetic is stood in by the oracle (the CUDA layer cannot
run here), so the summed result must equal the unpartitioned oracle, forward and backward."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn as nn

from oracle import reference_port as rp


def test_balanced_bounds():
    from mrgcn_b200.partition import balanced_bounds
    w = np.array([1, 1, 1, 1, 100, 1, 1, 1, 1, 1], dtype=float)
    b = balanced_bounds(w, 2)
    assert b[0] == 0 and b[-1] == 10 and np.all(np.diff(b) >= 0)
    assert abs(w[:b[1]].sum() - w[b[1]:].sum()) <= 104     # cuts are rounded to multiples of 4 nodes
    assert balanced_bounds(w, 2, align=1)[1] == 5
    b = balanced_bounds(np.ones(1000), 8)
    assert np.all(np.abs(np.diff(b) - 125) <= 4) and np.all(b[1:-1] % 4 == 0)
    assert balanced_bounds(np.ones(3), 8)[-1] == 3          # more ranks than nodes: empty ranges allowed


def test_split_coo_partitions_every_entry_once():
    from mrgcn_b200.partition import _Layout, balanced_bounds, node_weights, split_coo
    from mrgcn_b200.synth import synth_triples
    N, P = 300, 4
    R = 2 * P + 1
    A = rp.as_float32(rp.stacked_adjacency(synth_triples(N, P, 2000, seed=2), N, P)).tocoo()
    row, col, val = (torch.from_numpy(np.asarray(a)) for a in (A.row.astype(np.int64), A.col.astype(np.int64), A.data))
    bounds = balanced_bounds(node_weights(row, col, N), 3)
    nf = ni = 0
    seen = torch.zeros(N, dtype=torch.bool)
    for p in range(3):
        lay = _Layout(bounds, p)
        assert lay.maxrows % 4 == 0 and lay.NP == 3 * lay.maxrows >= N
        ids = torch.arange(lay.lo, lay.hi)
        assert torch.equal(lay.pad_ids(ids), p * lay.maxrows + torch.arange(lay.n_own))      # own rows: one contiguous block
        seen[ids] = True
        (fr, fc, fv), (ir, ic, iv) = split_coo(row, col, val, N, R, lay)
        nf += len(fr)
        ni += len(ir)
        assert fr.min() >= 0 and fr.max() < lay.n_own and fc.max() < R * lay.NP     # rows local, sources in the padded layout
        assert ir.max() < lay.NP and ic.max() < R * lay.n_own                        # rows in the padded layout, sources local
        x = torch.arange(N * 2, dtype=torch.float32).view(N, 2)
        assert torch.equal(lay.to_padded(x)[lay.own], x[lay.lo:lay.hi])
    assert nf == len(row) and ni == len(row) and bool(seen.all())


def _oracle_layer(X, weight_I, comp_I, weight_F, comp_F, bias, row_mask, gI, gF, B, relu, addend=None):
    """Stand-in for mrgcn_b200.layers.graph._LayerFn.apply with the oracle's arithmetic (graph.py:62-102)."""
    out = 0.0
    if weight_I is not None:
        W = weight_I
        if B > 0:
            ns = gI.shape[1] // comp_I.shape[0]
            W = torch.einsum("rb,bij->rij", comp_I, W.view(B, ns, -1)).reshape(comp_I.shape[0] * ns, -1)
        out = torch.mm(gI, W)
    if X is not None:
        W = weight_F if B <= 0 else torch.einsum("rb,bij->rij", comp_F, weight_F)
        R = W.shape[0]
        out = out + torch.mm(gF, torch.einsum("ij,bjk->bik", X, W).reshape(R * X.shape[0], -1))
    if addend is not None:
        out = out + addend
    if bias is not None:
        out = out + bias
    return torch.relu(out) if relu else out


def _coo_graph(row, col, val, nrows, ncols, R):
    return torch.sparse_coo_tensor(torch.stack([row, col]), val, (nrows, ncols))


def _worker(rank, world, port, tmp):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from mrgcn_b200.partition import PartitionedRGCN, _Layout, balanced_bounds, node_weights
        from mrgcn_b200.synth import synth_triples
        N, P = 200, 3
        R = 2 * P + 1
        A = rp.as_float32(rp.stacked_adjacency(synth_triples(N, P, 1500, seed=4), N, P))
        coo = A.tocoo()
        row, col, val = (torch.from_numpy(np.asarray(a)) for a in (coo.row.astype(np.int64), coo.col.astype(np.int64), coo.data))
        bounds = balanced_bounds(node_weights(row, col, N), world)
        lo, hi = int(bounds[rank]), int(bounds[rank + 1])
        # (dims, bases, featureless): 2-layer featureful NC (feature term of layer 0 destination-partitioned), the same with a
        # shape the projection path covers (layer 0 entirely source-partitioned, X sharded), 2-layer featureless NC without
        # bases, 1-layer LP-style encoder
        for dims, B, fl in (((5, 6, 3), 2, False), ((32, 4, 3), 4, False), ((0, 6, 3), -1, True), ((0, 8), 2, True)):
            torch.manual_seed(0)
            modules = [(dims[k], dims[k + 1], "mrgcn", nn.ReLU() if (k + 2 < len(dims) or len(dims) == 2) else None)
                       for k in range(len(dims) - 1)]
            acts = ["relu" if m[3] is not None else None for m in modules]
            layers, _ = rp.init_rgcn_params(modules, R, N, B, fl, True, False)
            X = None if fl else torch.randn(N, dims[0])
            G = torch.randn(N, dims[-1])
            for l in layers:
                for v in l.values():
                    v.requires_grad_(True)
            ref = rp.rgcn_forward(layers, acts, X, rp.csr_to_coo(A, torch.float32), num_nodes=N, num_relations=R,
                                  num_bases=B, featureless=fl)
            (ref * G).sum().backward()
            model = PartitionedRGCN(modules, R, N, B, fl, True, False, bounds, rank, layer_fn=_oracle_layer, graph_fn=_coo_graph)
            model.set_graph(row, col, val)
            # replicated parameters are broadcast from rank 0 when the graph is set: equal on every rank without any loading
            flat = torch.cat([p.detach().reshape(-1) for p in model.replicated_parameters()])
            both = [torch.empty_like(flat) for _ in range(world)]
            dist.all_gather(both, flat)
            assert all(torch.equal(both[0], b) for b in both)
            full_state = {"layers.layer_%d.%s" % (k, n): v.detach() for k, l in enumerate(layers) for n, v in l.items()}
            model.load_full_state(full_state)
            # state_dict() is in the reference layout again (weight_I gathered), so checkpoints are interchangeable
            sd = model.state_dict()
            for k_, v_ in full_state.items():
                assert torch.equal(sd[k_], v_), k_
            assert model.layer0_is_source_partitioned() == (fl or dims[0] == 32)
            X_in = None if fl else (X[lo:hi] if model.layer0_is_source_partitioned() else model.lay.to_padded(X))
            out = model(X_in)
            if len(dims) == 3:      # what a caller of the drop-in sees: all rows, true node order, on every rank
                with torch.no_grad():
                    assert torch.allclose(model.forward_all(X_in), ref.detach(), rtol=1e-5, atol=1e-6)
            assert out.shape == (hi - lo, dims[-1])
            assert torch.allclose(out, ref[lo:hi].detach(), rtol=1e-5, atol=1e-6), (dims, float((out - ref[lo:hi]).abs().max()))
            (out * G[lo:hi]).sum().backward()
            model.sync_grads()
            S = B if B > 0 else R
            for k, l in enumerate(layers):
                for n, v in l.items():
                    g = dict(model.named_parameters())["layers.layer_%d.%s" % (k, n)].grad
                    want = v.grad
                    if k == 0 and n == "weight_I":
                        want = want.view(S, N, -1)[:, lo:hi, :].reshape(S * (hi - lo), -1)
                    assert torch.allclose(g, want, rtol=1e-4, atol=1e-5), (dims, k, n, float((g - want).abs().max()))
        open(os.path.join(tmp, "ok%d" % rank), "w").write("ok")
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_partitioned_rgcn_two_ranks_gloo(tmp_path):
    port = 29500 + (os.getpid() % 400)
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert os.path.exists(tmp_path / "ok0") and os.path.exists(tmp_path / "ok1")

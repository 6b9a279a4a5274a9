#!/usr/bin/env python
"""Generate golden input/output vectors by running the UNMODIFIED reference.

Run in the build container only (needs /root/reference, which is absent on the GPU box):

    python tests/golden/make_golden.py

It imports wxwilcke/mrgcn from /root/reference (with the import-only rdflib stand-in under
tests/golden/_stubs, because mrgcn/data/utils.py:10 imports rdflib at module scope), drives the
reference's own classes on small seeded inputs and writes inputs + outputs to
tests/golden/*.npz.  Those files pin `oracle/reference_port.py` (tests/test_oracle_golden.py)
and are a second, reference-made yardstick for the CUDA path (tests/test_gpu_*.py).
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "_stubs"))
sys.path.insert(0, "/root/reference")

import numpy as np  # noqa: E402
import scipy.sparse as sp  # noqa: E402
import torch  # noqa: E402
import torch.nn as nn  # noqa: E402

from mrgcn.data.batch import A_Batch, FullBatch, MiniBatch  # noqa: E402
from mrgcn.encodings import graph_structure  # noqa: E402
from mrgcn.layers.graph import GraphConvolution  # noqa: E402
from mrgcn.models.mrgcn import MRGCN  # noqa: E402
from mrgcn.models.rgcn import RGCN  # noqa: E402
from mrgcn.tasks import link_prediction as lp  # noqa: E402

sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from mrgcn_b200.synth import synth_triples  # noqa: E402


class FakeKG:
    """Just enough of mrgcn.data.io.knowledge_graph.KnowledgeGraph for graph_structure.generate:
    nodes are named so that sorting by str() equals sorting by integer id."""

    def __init__(self, triples, num_nodes):
        self.t = [("n%07d" % s, "p%04d" % p, "n%07d" % o) for s, p, o in triples.tolist()]
        self.n = num_nodes

    def properties(self):
        return (p for _, p, _ in self.t)

    def atoms(self, separate_literals=True):
        return {"n%07d" % i for i in range(self.n)}

    def quickSort(self, lst):
        return sorted(lst, key=str)

    def property_frequency(self, prop):
        return sum(1 for _, p, _ in self.t if p == prop)

    def triples(self, triple=(None, None, None), separate_literals=True):
        _, prop, _ = triple
        return (x for x in self.t if prop is None or x[1] == prop)


def reference_adjacency(triples, num_nodes):
    cfg = {"graph": {"structural": {"separate_literals": True, "include_inverse_properties": True,
                                    "exclude_properties": [], "multiprocessing": False}}}
    A, nodes, props = graph_structure.generate(FakeKG(triples, num_nodes), cfg)
    assert [nodes["n%07d" % i] for i in range(num_nodes)] == list(range(num_nodes))
    return A


def f32(A):
    # mrgcn/data/io/tarball.py:151-157
    return sp.csr_matrix((A.data.astype(np.float32), A.indices, A.indptr), shape=A.shape)


def grads(params):
    return {"grad_" + k: v.grad.detach().numpy().copy() for k, v in params if v.grad is not None}


def save(name, **arrs):
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **arrs)
    print("wrote", path, os.path.getsize(path), "bytes")


def case_adjacency():
    N, P = 61, 4
    tr = synth_triples(N, P, 150, seed=3)
    A = reference_adjacency(tr, N)
    Af = f32(A)
    fb = FullBatch(Af, [np.empty((N, 0), dtype=np.float32)], np.arange(N))
    fb.as_tensors_()
    save("adjacency", triples=tr, num_nodes=N, num_props=P,
         data64=A.data, data32=Af.data, indices=A.indices, indptr=A.indptr, shape=np.array(A.shape),
         coo_indices=fb.A._indices().numpy(), coo_values_int8=fb.A._values().numpy())
    return tr, N, P, Af


def layer_case(name, Af, N, R, *, indim, outdim, num_bases, bias, input_layer, featureless,
               int8_values, seed):
    torch.manual_seed(seed)
    layer = GraphConvolution(indim, outdim, R, N, num_bases=num_bases, bias=bias,
                             input_layer=input_layer, featureless=featureless)
    if bias:
        with torch.no_grad():
            layer.b.uniform_(-0.5, 0.5)
    idx = torch.LongTensor(np.array(Af.nonzero()))
    vals = torch.Tensor(Af.data)
    A = torch.sparse_coo_tensor(idx, vals, Af.shape, dtype=torch.int8 if int8_values else torch.float32)
    X = None
    if not (input_layer and featureless):
        X = torch.randn(N, indim, requires_grad=True)
    out = layer(X, A)
    G = torch.randn_like(out)
    (out * G).sum().backward()
    arrs = {"param_" + k: v.detach().numpy().copy() for k, v in layer.named_parameters()}
    arrs.update(grads(layer.named_parameters()))
    if X is not None:
        arrs["X"] = X.detach().numpy()
        arrs["grad_X"] = X.grad.numpy()
    save(name, out=out.detach().numpy(), G=G.numpy(), a_values=A._values().numpy(),
         a_indices=A._indices().numpy(),
         meta=np.array([indim, outdim, R, N, num_bases, int(bias), int(input_layer), int(featureless)]),
         **arrs)


def case_rgcn(tr, N, P, Af):
    R = 2 * P + 1
    torch.manual_seed(11)
    X = torch.randn(N, 7)
    for tag, bases, featureless, lpred in (("rgcn_nc_basis", 3, False, False), ("rgcn_nc_featureless", -1, True, False),
                                           ("rgcn_lp_basis", 2, True, True)):
        torch.manual_seed(5)
        if lpred:
            modules = [(0, 12, "mrgcn", nn.ReLU())]
        else:
            modules = [(0 if featureless else 7, 6, "mrgcn", nn.ReLU()), (6, 3, "mrgcn", None)]
        model = MRGCN(modules, [], R, N, num_bases=bases, p_dropout=0.0, featureless=featureless,
                      bias=True, link_prediction=lpred)
        fb = FullBatch(Af, [X.numpy().copy() if not featureless else np.empty((N, 0), dtype=np.float32)],
                       np.arange(N))
        fb.as_tensors_()
        # float-valued A (true 1/deg): the layer's A.float() is then a no-op
        fb.A = torch.sparse_coo_tensor(fb.A._indices(), torch.Tensor(Af.data), Af.shape)
        if featureless:
            out = model(fb)
        else:
            out = model.rgcn(fb.X[0].float(), fb.A)
        arrs = {"param_" + k: v.detach().numpy().copy() for k, v in model.named_parameters()}
        if lpred:
            data = torch.as_tensor(tr[:40].astype(np.int64))
            np.random.seed(123)
            # link_prediction.py:244-268 (negative sampling) re-run verbatim through the reference
            # by calling the same numpy calls in the same order
            n = data.shape[0]
            nodes = np.union1d(data[:, 0], data[:, 2])
            nc = n // 5
            pick = np.random.choice(np.arange(n), nc, replace=False)
            nh = nc // 2
            nt = nc - nh
            corrupted = np.empty((nc, 3), dtype=int)
            corrupted[:] = data.numpy()[pick]
            corrupted[:nh, 0] = np.random.choice(nodes, nh)
            corrupted[-nt:, 2] = np.random.choice(nodes, nt)
            cd = torch.as_tensor(corrupted).long()
            Y = torch.ones(n + nc)
            Y[-nc:] = 0
            Yh = torch.empty(n + nc)
            Yh[:n] = lp.score_distmult_bc((data[:, 0], data[:, 1], data[:, 2]), out, model.rgcn.relations)
            Yh[-nc:] = lp.score_distmult_bc((cd[:, 0], cd[:, 1], cd[:, 2]), out, model.rgcn.relations)
            loss = lp.binary_crossentropy(Yh, Y, nn.BCEWithLogitsLoss())
            loss.backward()
            with torch.no_grad():
                emb = model(fb)
                ranks_raw = lp.compute_ranks_fast(data, emb, model.rgcn.relations, 16, False)
                ranks_flt = lp.compute_ranks_fast(data, emb, model.rgcn.relations, 16, True)
            arrs.update(data=data.numpy(), corrupted=corrupted, scores=Yh.detach().numpy(),
                        ranks_raw=ranks_raw.numpy(), ranks_flt=ranks_flt.numpy(), emb=emb.numpy())
        else:
            labelled = torch.arange(0, N, 3)
            targets = (labelled * 7) % 3
            loss = nn.CrossEntropyLoss()(out[labelled], targets)
            loss.backward()
            arrs.update(labelled=labelled.numpy(), targets=targets.numpy())
        arrs.update(grads(model.named_parameters()))
        save(tag, out=out.detach().numpy(), loss=np.array(loss.item()), X=X.numpy(),
             meta=np.array([R, N, bases, int(featureless), int(lpred)]), **arrs)


def case_minibatch(tr, N, P, Af):
    """rgcn.py:91-128 + batch.py:168-263: two-layer mini-batch forward/backward."""
    R = 2 * P + 1
    torch.manual_seed(21)
    X = torch.randn(N, 5)
    model = RGCN([(5, 6, "mrgcn", nn.ReLU()), (6, 3, "mrgcn", None)], R, N, 2, 0.0, False, True, False)
    batch_idx = np.array([2, 3, 11, 40])
    ab = A_Batch(Af, batch_idx, 2)
    outer = ab.neighbours[-1].copy()
    ab.as_tensors_()
    Xo = X[torch.from_numpy(outer)].clone().requires_grad_(True)
    out = model(Xo, ab)
    G = torch.randn_like(out)
    (out * G).sum().backward()
    arrs = {"param_" + k: v.detach().numpy().copy() for k, v in model.named_parameters()}
    arrs.update(grads(model.named_parameters()))
    save("rgcn_minibatch", out=out.detach().numpy(), G=G.numpy(), X=X.numpy(), batch_idx=batch_idx,
         neigh0=ab.neighbours[0].numpy(), neigh1=ab.neighbours[1].numpy(),
         row0_idx=ab.row[0]._indices().numpy(), row0_val=ab.row[0]._values().numpy(),
         row1_idx=ab.row[1]._indices().numpy(), row1_val=ab.row[1]._values().numpy(),
         grad_X=Xo.grad.numpy(), meta=np.array([R, N, 2]), **arrs)


def case_mrgcn_modalities(tr, N, P, Af):
    """MRGCN.forward with gated literal encoders (mrgcn.py:189-214,250-305): numeric MLP + temporal MLP whose
    outputs are scaled by gate_weights and scattered into the rows of the nodes that carry the literal."""
    R = 2 * P + 1
    torch.manual_seed(31)
    rng = np.random.default_rng(31)
    idx_num = np.sort(rng.choice(N, 23, replace=False))
    idx_date = np.sort(rng.choice(N, 17, replace=False))
    enc_num = rng.normal(size=(23, 3)).astype(np.float32)
    enc_date = rng.normal(size=(17, 6)).astype(np.float32)
    emb = [("xsd.numeric", (3, 2, 0.0), False), ("xsd.date", (6, 4, 0.0), False)]
    modules = [(6, 5, "mrgcn", nn.ReLU()), (5, 3, "mrgcn", None)]
    model = MRGCN(modules, emb, R, N, num_bases=2, p_dropout=0.0, featureless=False, bias=True)
    X = [np.empty((N, 0), dtype=np.float32),
         ["xsd.numeric", [[enc_num, idx_num, np.full(23, -1)]], False],
         ["xsd.date", [[enc_date, idx_date, np.full(17, -1)]], False]]
    fb = FullBatch(Af, X, np.arange(N))
    fb.as_tensors_()
    fb.A = torch.sparse_coo_tensor(fb.A._indices(), torch.Tensor(Af.data), Af.shape)
    out = model(fb)
    G = torch.randn_like(out)
    (out * G).sum().backward()
    arrs = {"param_" + k: v.detach().numpy().copy() for k, v in model.named_parameters()}
    arrs.update(grads(model.named_parameters()))
    save("mrgcn_modalities", out=out.detach().numpy(), G=G.numpy(), idx_num=idx_num, idx_date=idx_date, enc_num=enc_num,
         enc_date=enc_date, meta=np.array([R, N, 2]), **arrs)


def case_distmult():
    torch.manual_seed(9)
    E = torch.randn(50, 24, requires_grad=True)
    Rel = torch.randn(7, 24, requires_grad=True)
    g = torch.Generator().manual_seed(4)
    s = torch.randint(0, 50, (90,), generator=g)
    p = torch.randint(0, 7, (90,), generator=g)
    o = torch.randint(0, 50, (90,), generator=g)
    sc = lp.score_distmult_bc((s, p, o), E, Rel)
    G = torch.randn_like(sc)
    (sc * G).sum().backward()
    # the broadcast short-cuts (link_prediction.py:652-663)
    with torch.no_grad():
        sb = lp.score_distmult_bc((torch.arange(50).view(1, 50, 1).expand(4, 50, 1), p[:4].view(4, 1, 1),
                                   o[:4].view(4, 1, 1)), E, Rel)
    save("distmult", E=E.detach().numpy(), Rel=Rel.detach().numpy(), s=s.numpy(), p=p.numpy(), o=o.numpy(),
         scores=sc.detach().numpy(), G=G.numpy(), grad_E=E.grad.numpy(), grad_Rel=Rel.grad.numpy(),
         scores_bc=sb.numpy())


def main():
    tr, N, P, Af = case_adjacency()
    R = 2 * P + 1
    k = 0
    for int8 in (False, True):
        sfx = "_int8" if int8 else ""
        for nb in (-1, 3):
            b = "_b%d" % nb if nb > 0 else ""
            layer_case("layer_input_featureless%s%s" % (b, sfx), Af, N, R, indim=0, outdim=5, num_bases=nb,
                       bias=True, input_layer=True, featureless=True, int8_values=int8, seed=100 + k)
            layer_case("layer_input_features%s%s" % (b, sfx), Af, N, R, indim=9, outdim=6, num_bases=nb,
                       bias=(nb > 0), input_layer=True, featureless=False, int8_values=int8, seed=200 + k)
            layer_case("layer_hidden%s%s" % (b, sfx), Af, N, R, indim=6, outdim=4, num_bases=nb,
                       bias=True, input_layer=False, featureless=False, int8_values=int8, seed=300 + k)
            k += 1
    case_rgcn(tr, N, P, Af)
    case_minibatch(tr, N, P, Af)
    case_mrgcn_modalities(tr, N, P, Af)
    case_distmult()


if __name__ == "__main__":
    main()

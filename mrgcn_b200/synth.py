"""Synthetic knowledge graphs of the shapes BASELINE.json names.

Host-side NumPy only; produces *integer triples* (s, p, o).  The stacked adjacency itself is
built on the GPU by `mrgcn_b200.graph.RelGraph.from_triples` (product path) or, in tests, by
`oracle.reference_port.stacked_adjacency` (the reference's scipy recipe,
/root/reference/mrgcn/encodings/graph_structure.py:70-108).
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np


@dataclass(frozen=True)
class Shape:
    name: str
    num_nodes: int
    num_props: int          # P; relations R = 2P + 1 (inverse + self-loop block)
    num_triples: int        # T; nnz(A) = 2T + N
    num_bases: int
    dims: tuple             # (in0, h1, ..., out); in0 == 0 -> featureless input layer
    task: str               # "nc" | "lp"

    @property
    def num_relations(self):
        return 2 * self.num_props + 1

    @property
    def nnz(self):
        return 2 * self.num_triples + self.num_nodes


# config shapes (from BASELINE.json `configs` + the reference TOMLs)
SHAPES = {
    "synth": Shape("synth", 2329, 14, 2594, 0, (0, 16, 2), "nc"),
    "aifb": Shape("aifb", 8285, 45, 29043, 0, (0, 16, 4), "nc"),
    "aifb_b40": Shape("aifb_b40", 8285, 45, 29043, 40, (0, 16, 4), "nc"),
    "am": Shape("am", 1666764, 133, 5900000, 40, (151, 10, 11), "nc"),
    "am16": Shape("am16", 104172, 133, 368750, 40, (151, 10, 11), "nc"),
    "am8": Shape("am8", 208345, 133, 737500, 40, (151, 10, 11), "nc"),
    "fb15k237": Shape("fb15k237", 14541, 237, 310116, 2, (0, 200), "lp"),
    "yago3-10+": Shape("yago3-10+", 250000, 45, 1090000, 2, (145, 200), "lp"),
}


def synth_triples(num_nodes, num_props, num_triples, seed=0, node_dist="powerlaw"):
    """Draw unique triples: p ~ truncated Zipf(1.0) (every property gets >= 1 triple),
    s, o ~ power law (id = floor(N * u^2), then a fixed random permutation) or uniform.
    Returns an int32 array (T', 3), T' <= num_triples after de-duplication (RDF graphs are
    sets), in a seeded random order."""
    rng = np.random.default_rng(seed)
    w = 1.0 / np.arange(1, num_props + 1)
    p = rng.choice(num_props, size=num_triples, p=w / w.sum())
    p[:num_props] = np.arange(num_props)
    if node_dist == "powerlaw":
        perm = rng.permutation(num_nodes)
        s = perm[np.minimum((num_nodes * rng.random(num_triples) ** 2).astype(np.int64), num_nodes - 1)]
        o = perm[np.minimum((num_nodes * rng.random(num_triples) ** 2).astype(np.int64), num_nodes - 1)]
    elif node_dist == "uniform":
        s = rng.integers(0, num_nodes, num_triples)
        o = rng.integers(0, num_nodes, num_triples)
    else:
        raise ValueError(node_dist)
    key = (s.astype(np.int64) * num_props + p) * num_nodes + o
    _, first = np.unique(key, return_index=True)
    first.sort()
    return np.stack([s[first], p[first], o[first]], axis=1).astype(np.int32)


def synth_graph(shape: Shape | str, seed=0, node_dist="powerlaw", scale=1.0):
    """Triples for a named shape; `scale` < 1 shrinks N and T together (R, B, dims unchanged)."""
    if isinstance(shape, str):
        shape = SHAPES[shape]
    n = max(int(shape.num_nodes * scale), 4)
    t = max(int(shape.num_triples * scale), shape.num_props)
    return n, synth_triples(n, shape.num_props, t, seed, node_dist)

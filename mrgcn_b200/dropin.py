"""Swap the reference's hot-path classes and functions for the B200 ones at their import paths, so that
`mrgcn/run.py`, the task loops and the TOML configs run unchanged.

    import mrgcn_b200.dropin; mrgcn_b200.dropin.install()      # before `from mrgcn.tasks import ...` is used

What is replaced (reference path -> ours):
    mrgcn.layers.graph.GraphConvolution          -> mrgcn_b200.layers.graph.GraphConvolution
    mrgcn.models.rgcn.RGCN / .GraphConvolution   -> mrgcn_b200.models.rgcn.RGCN
    mrgcn.models.mrgcn.MRGCN / .RGCN             -> mrgcn_b200.models.mrgcn.MRGCN
    mrgcn.tasks.{node_classification,link_prediction}.MRGCN
    mrgcn.tasks.link_prediction.score_distmult_bc / compute_ranks_fast
The reference's batch classes stay: our layer accepts the CPU sparse COO tensor they hand over
(`Batch.to` never moves a FullBatch's A, mrgcn/data/batch.py:122-123) and caches the device graph on it; our
MRGCN recognises the reference's MiniBatch / A_Batch by name.
"""
from __future__ import annotations

import importlib


def install():
    from .layers.graph import GraphConvolution
    from .models.mrgcn import MRGCN
    from .models.rgcn import RGCN
    from .tasks import link_prediction as lp_b200

    graph = importlib.import_module("mrgcn.layers.graph")
    rgcn = importlib.import_module("mrgcn.models.rgcn")
    mrgcn = importlib.import_module("mrgcn.models.mrgcn")
    graph.GraphConvolution = GraphConvolution
    rgcn.GraphConvolution = GraphConvolution
    rgcn.RGCN = RGCN
    mrgcn.RGCN = RGCN
    mrgcn.MRGCN = MRGCN
    for name in ("mrgcn.tasks.node_classification", "mrgcn.tasks.link_prediction"):
        try:
            mod = importlib.import_module(name)
        except ImportError:        # a task module's own dependencies (rdflib, ...) may be absent
            continue
        mod.MRGCN = MRGCN
        if name.endswith("link_prediction"):
            mod.score_distmult_bc = lp_b200.score_distmult_bc
            mod.compute_ranks_fast = lp_b200.compute_ranks_fast
    return True

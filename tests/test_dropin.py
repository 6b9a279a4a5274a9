"""dropin.install() rebinds the reference's import paths (needs /root/reference: build container only)."""
import os
import sys

import pytest

REF = "/root/reference"
STUBS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "_stubs")


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not present (GPU box)")
def test_install_rebinds_reference_paths():
    sys.path.insert(0, STUBS)
    sys.path.insert(0, REF)
    try:
        import mrgcn_b200.dropin as dropin
        assert dropin.install()
        import mrgcn.layers.graph as g
        import mrgcn.models.mrgcn as m
        import mrgcn.models.rgcn as r
        import mrgcn.tasks.link_prediction as lp
        import mrgcn.tasks.node_classification as nc
        from mrgcn_b200.layers.graph import GraphConvolution
        from mrgcn_b200.models.mrgcn import MRGCN
        from mrgcn_b200.models.rgcn import RGCN
        from mrgcn_b200.tasks import link_prediction as ours
        assert g.GraphConvolution is GraphConvolution and r.GraphConvolution is GraphConvolution
        assert r.RGCN is RGCN and m.RGCN is RGCN and m.MRGCN is MRGCN
        assert nc.MRGCN is MRGCN and lp.MRGCN is MRGCN
        assert lp.score_distmult_bc is ours.score_distmult_bc and lp.compute_ranks_fast is ours.compute_ranks_fast
        # the reference's own model builder now yields our classes with the reference's state_dict keys
        import torch.nn as nn
        model = m.MRGCN([(0, 4, "mrgcn", nn.ReLU()), (4, 2, "mrgcn", None)], [], 5, 7, num_bases=2, featureless=True, bias=True)
        assert list(model.state_dict().keys()) == [
            "gate_weights"][:0] + ["rgcn.layers.layer_0.weight_I_comp", "rgcn.layers.layer_0.weight_I", "rgcn.layers.layer_0.b",
                                   "rgcn.layers.layer_1.weight_F_comp", "rgcn.layers.layer_1.weight_F", "rgcn.layers.layer_1.b"]
    finally:
        for p in (REF, STUBS):
            if p in sys.path:
                sys.path.remove(p)
        for k in [k for k in sys.modules if k == "mrgcn" or k.startswith("mrgcn.") or k.startswith("rdflib")]:
            del sys.modules[k]

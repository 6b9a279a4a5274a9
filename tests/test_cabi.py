"""The C-ABI library loads and exports every symbol include/mrgcn_b200.h declares, and the ctypes mirror of its
structs has the layout the C compiler gives them.  No compute calls (CPU only)."""
import ctypes
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "mrgcn_b200.h")


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mrgcn_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from mrgcn_b200 import _native
    assert os.path.exists(_native.LIB_PATH), "build first: python -c 'import __graft_entry__ as g; g.build()'"
    lib = ctypes.CDLL(_native.LIB_PATH)
    names = declared_functions()
    assert len(names) >= 12
    for n in names:
        assert hasattr(lib, n), "libmrgcn_b200.so does not export %s" % n
    assert set(names) == set(_native.SYMBOLS), "ctypes table and header disagree: %s" % (set(names) ^ set(_native.SYMBOLS))
    assert _native.lib().mrgcn_version() >= 100


def test_ctypes_struct_layout_matches_c(tmp_path):
    from mrgcn_b200 import _native
    prog = tmp_path / "layout.c"
    prog.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "%s"\n'
                    'int main(void){printf("%%zu %%zu %%zu %%zu %%zu %%zu %%zu %%zu %%zu %%zu %%zu\\n", sizeof(mrgcn_graph), sizeof(mrgcn_layer_args),'
                    ' sizeof(mrgcn_layer_bwd_args), offsetof(mrgcn_graph, long_rows), offsetof(mrgcn_graph, n_chunks),'
                    ' offsetof(mrgcn_layer_args, addend), offsetof(mrgcn_layer_bwd_args, gact), sizeof(mrgcn_tab_plan),'
                    ' offsetof(mrgcn_tab_plan, n_blks), offsetof(mrgcn_layer_args, x_stride), offsetof(mrgcn_layer_bwd_args, phases));return 0;}\n' % HEADER)
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-o", str(exe), str(prog)], check=True)
    got = [int(x) for x in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()]
    want = [ctypes.sizeof(_native.Graph), ctypes.sizeof(_native.LayerArgs), ctypes.sizeof(_native.LayerBwdArgs),
            _native.Graph.long_rows.offset, _native.Graph.n_chunks.offset, _native.LayerArgs.addend.offset,
            _native.LayerBwdArgs.gact.offset, ctypes.sizeof(_native.TabPlan), _native.TabPlan.n_blks.offset,
            _native.LayerArgs.x_stride.offset, _native.LayerBwdArgs.phases.offset]
    assert got == want


def test_product_path_refuses_cpu():
    """No CPU fallback: the layer raises instead of computing on the host."""
    import torch
    from mrgcn_b200.layers.graph import GraphConvolution
    layer = GraphConvolution(3, 2, 3, 4, input_layer=False)
    A = torch.sparse_coo_tensor(torch.tensor([[0, 1], [1, 2]]), torch.ones(2), (4, 12))
    with pytest.raises(RuntimeError, match="no CPU path"):
        layer(torch.randn(4, 3), A)


def test_product_does_not_import_oracle():
    out = subprocess.run([sys.executable, "-c",
                          "import sys; sys.path.insert(0, %r); import mrgcn_b200.layers.graph, mrgcn_b200.models.mrgcn, "
                          "mrgcn_b200.tasks.link_prediction, mrgcn_b200.partition; "
                          "print(any(m.startswith('oracle') for m in sys.modules))" % ROOT],
                         check=True, capture_output=True, text=True).stdout.strip()
    assert out == "False"


def test_layer_init_matches_reference(golden):
    """Same shapes, registration order and initialisers as the reference layer: the same seed gives the same
    initial state_dict (graph.py:9-60,104-116).  Construction works without a GPU."""
    import torch
    from mrgcn_b200.layers.graph import GraphConvolution
    g = golden("layer_input_features_b3")
    indim, outdim, R, N, nb, bias, inp, fl = (int(v) for v in g["meta"])
    torch.manual_seed(201)          # tests/golden/make_golden.py: seed 200 + k, k = 1 for (float values, 3 bases)
    layer = GraphConvolution(indim, outdim, R, N, num_bases=nb, bias=bool(bias), input_layer=bool(inp), featureless=bool(fl))
    assert [k for k, _ in layer.named_parameters()] == ["weight_I_comp", "weight_F_comp", "weight_I", "weight_F", "b"]
    for k in ("weight_I_comp", "weight_F_comp", "weight_I", "weight_F"):
        assert torch.equal(getattr(layer, k).detach(), torch.from_numpy(g["param_" + k])), k


def test_model_state_dict_keys_match_reference_layout():
    """Checkpoints are plain state_dicts (run.py:230-236): names, order and shapes must be the reference's."""
    import torch.nn as nn
    from mrgcn_b200.models.mrgcn import MRGCN
    m = MRGCN([(7, 6, "mrgcn", nn.ReLU()), (6, 3, "mrgcn", None)], [("xsd.numeric", (3, 2, 0.0), False)], 9, 61, num_bases=3,
              featureless=False, bias=True, link_prediction=True)
    keys = list(m.state_dict().keys())
    assert keys == ["gate_weights", "module_dict.xsd_numeric_0.mlp.0.weight", "module_dict.xsd_numeric_0.mlp.0.bias",
                    "rgcn.relations",
                    "rgcn.layers.layer_0.weight_I_comp", "rgcn.layers.layer_0.weight_F_comp", "rgcn.layers.layer_0.weight_I",
                    "rgcn.layers.layer_0.weight_F", "rgcn.layers.layer_0.b",
                    "rgcn.layers.layer_1.weight_F_comp", "rgcn.layers.layer_1.weight_F", "rgcn.layers.layer_1.b"]
    sd = m.state_dict()
    assert tuple(sd["rgcn.layers.layer_0.weight_I"].shape) == (3 * 61, 6) and tuple(sd["rgcn.relations"].shape) == (9, 3)
    assert m.devices["relational"].type in ("cuda", "cpu") and m.gate_map == {"xsd_numeric_0": 0}


def test_distmult_workspace_covers_index_lists_and_sort_storage():
    """mrgcn_distmult_bwd_ws_elems(n): 12 n index words + the radix sort's temporary storage (a size query: no GPU needed);
    nothing is allocated inside mrgcn_distmult_bwd (include/mrgcn_b200.h)."""
    from mrgcn_b200 import _native as nv
    lib = nv.lib()
    for n in (1, 600, 100000):
        w = int(lib.mrgcn_distmult_bwd_ws_elems(n))
        assert w >= 12 * n + 64, (n, w)
    assert int(lib.mrgcn_distmult_bwd_ws_elems(0)) == int(lib.mrgcn_distmult_bwd_ws_elems(1))

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
                    'int main(void){printf("%%zu %%zu %%zu %%zu %%zu %%zu %%zu\\n", sizeof(mrgcn_graph), sizeof(mrgcn_layer_args),'
                    ' sizeof(mrgcn_layer_bwd_args), offsetof(mrgcn_graph, long_rows), offsetof(mrgcn_graph, n_chunks),'
                    ' offsetof(mrgcn_layer_args, addend), offsetof(mrgcn_layer_bwd_args, gact));return 0;}\n' % HEADER)
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-o", str(exe), str(prog)], check=True)
    got = [int(x) for x in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()]
    want = [ctypes.sizeof(_native.Graph), ctypes.sizeof(_native.LayerArgs), ctypes.sizeof(_native.LayerBwdArgs),
            _native.Graph.long_rows.offset, _native.Graph.n_chunks.offset, _native.LayerArgs.addend.offset,
            _native.LayerBwdArgs.gact.offset]
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

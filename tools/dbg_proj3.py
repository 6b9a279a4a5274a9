import sys, os
sys.path.insert(0, os.getcwd())
import torch
from mrgcn_b200 import _native as nv
from mrgcn_b200.layers.graph import padded_features
DEV = "cuda"
torch.manual_seed(0)
N, indim, B, outdim = 20000, 151, 40, 10
X = torch.randn(N, indim); V = torch.randn(B, indim, outdim) * 0.1
Xd = padded_features(X.to(DEV)); Vd = V.to(DEV)
tru = torch.einsum("ij,bjk->ibk", X.double(), V.double()).reshape(N, B * outdim)
P = torch.full((N, B * outdim), float("nan"), device=DEV)
vt = torch.empty(2 * B * outdim * 160, device=DEV)
for rep in range(3):
    nv.check(nv.lib().mrgcn_feat_proj(nv.ptr(Xd), N, indim, Xd.stride(0), nv.ptr(Vd), B, outdim, nv.ptr(vt), None, nv.ptr(P), nv.stream_ptr()), "feat_proj")
    torch.cuda.synchronize()
    e = (P.cpu().double() - tru).abs()
    print("dbg", os.environ.get("MRGCN_PROJ_DEBUG"), "bad fraction", float((e > 1e-4).float().mean()), "max", float(e.max()))

import sys, os
sys.path.insert(0, os.getcwd())
import torch
from mrgcn_b200 import _native as nv
from mrgcn_b200.layers.graph import padded_features
DEV = "cuda"
def tf32(x):
    u = x.view(torch.int32)
    return ((u + 0x1000) & ~0x1FFF).view(torch.float32)
def run(X, V):
    N, indim = X.shape; B, _, outdim = V.shape
    pitch = int(nv.lib().mrgcn_feat_proj_supported(indim, B, outdim))
    Xd = padded_features(X.to(DEV)); Vd = V.to(DEV)
    P = torch.full((N, B * outdim), float("nan"), device=DEV)
    vt = torch.empty(2 * B * outdim * pitch, device=DEV)
    nv.check(nv.lib().mrgcn_feat_proj(nv.ptr(Xd), N, indim, Xd.stride(0), nv.ptr(Vd), B, outdim, nv.ptr(vt), None, nv.ptr(P), nv.stream_ptr()), "feat_proj")
    torch.cuda.synchronize()
    tru = torch.einsum("ij,bjk->ibk", X.double(), V.double()).reshape(N, B * outdim)
    e = (P.cpu().double() - tru).abs()
    return float(e.max()), float(e.pow(2).mean().sqrt()), float(tru.abs().max())
torch.manual_seed(0)
for (N, indim, B, outdim) in [(256, 64, 8, 16), (1000, 151, 40, 10)]:
    X = torch.randn(N, indim); V = torch.randn(B, indim, outdim) * 0.1
    print(N, indim, B, outdim)
    print("  X tf32, V tf32  :", run(tf32(X), tf32(V)))
    print("  X full, V tf32  :", run(X, tf32(V)))
    print("  X tf32, V full  :", run(tf32(X), V))
    print("  X full, V full  :", run(X, V))
    # one-hot probes: X = e_k rows -> P rows = V[:, k, :] exactly
    Xo = torch.zeros(N, indim); Xo[torch.arange(N), torch.arange(N) % indim] = 1.0
    print("  one-hot X       :", run(Xo, V))
    Xo2 = Xo * 1.0001220703125  # 1 + 2^-13: hi = 1, lo = 2^-13
    print("  one-hot(1+2^-13):", run(Xo2, tf32(V)))

print("multi-item:")
for N in (20000, 208345):
    X = torch.randn(N, 151); V = torch.randn(40, 151, 10) * 0.1
    print(" ", N, "X full, V full  :", run(X, V))
# upload path
import ctypes as C
X = torch.randn(50000, 151)
buf = torch.zeros((50000, 160), device=DEV)
nv.check(nv.lib().mrgcn_upload_rows(X.data_ptr(), 50000, 151, buf.data_ptr(), 160, nv.stream_ptr()), "upload")
torch.cuda.synchronize()
print("upload pageable ok:", torch.equal(buf[:, :151].cpu(), X), float(buf[:, 151:].abs().max()))
Xp = X.pin_memory()
buf.zero_()
nv.check(nv.lib().mrgcn_upload_rows(Xp.data_ptr(), 50000, 151, buf.data_ptr(), 160, nv.stream_ptr()), "upload")
torch.cuda.synchronize()
print("upload pinned ok:", torch.equal(buf[:, :151].cpu(), X))

import sys, os
sys.path.insert(0, os.getcwd())
import torch
from mrgcn_b200 import _native as nv
from mrgcn_b200.layers.graph import padded_features
DEV = "cuda"
torch.manual_seed(0)
N, indim, B, outdim = 20000, 151, 40, 10
X = torch.randn(N, indim); V = torch.randn(B, indim, outdim) * 0.1
pitch = 160
Xd = padded_features(X.to(DEV)); Vd = V.to(DEV)
P = torch.full((N, B * outdim), float("nan"), device=DEV)
vt = torch.empty(2 * B * outdim * pitch, device=DEV)
nv.check(nv.lib().mrgcn_feat_proj(nv.ptr(Xd), N, indim, Xd.stride(0), nv.ptr(Vd), B, outdim, nv.ptr(vt), None, nv.ptr(P), nv.stream_ptr()), "feat_proj")
torch.cuda.synchronize()
tru = torch.einsum("ij,bjk->ibk", X.double(), V.double()).reshape(N, B * outdim)
e = (P.cpu().double() - tru).abs()
bad = e > 1e-4
print("bad fraction", float(bad.float().mean()), "nan", int(torch.isnan(P).sum()))
nt = (N + 127) // 128
GR = 29
for t in range(nt):
    rows = slice(t * 128, min(N, (t + 1) * 128))
    per_chunk = [float(bad[rows, c * 80:(c + 1) * 80].float().mean()) for c in range(5)]
    if max(per_chunk) > 0:
        # which row quarters are bad
        q = [float(bad[t * 128 + 32 * k: t * 128 + 32 * (k + 1)].float().mean()) for k in range(4)]
        print("tile %3d (rg %2d, it %d) bad per chunk %s per row-quarter %s" % (t, t % GR, t // GR, ["%.2f" % x for x in per_chunk], ["%.2f" % x for x in q]))
# does a bad row equal the product of some OTHER X row?  check tile t against tile t - GR etc.
Pc = P.cpu().double()
for t in range(nt):
    r0 = t * 128
    if bad[r0:r0 + 128].any():
        for dt in (-GR, GR, -1, 1):
            t2 = t + dt
            if 0 <= t2 < nt and t2 * 128 + 128 <= N:
                d = (Pc[r0:r0 + 128] - tru[t2 * 128:t2 * 128 + 128]).abs()
                if float((d < 1e-4).float().mean()) > 0.1:
                    print("tile", t, "matches truth of tile", t2, float((d < 1e-4).float().mean()))
        break
# column-wise structure of error in first bad tile
for t in range(nt):
    r0 = t * 128
    if bad[r0:r0 + 128].any():
        bb = bad[r0:r0 + 128]
        print("first bad tile", t, "bad cols (first 40):", bb.any(0)[:40].int().tolist())
        print("row 0 err", e[r0, :12].tolist())
        # hypothesis: missing / duplicated k-chunk contribution: compare with partial sums
        Xt = X[r0:r0 + 128].double()
        for kc in range(5):
            part = torch.einsum("ij,bjk->ibk", Xt[:, kc * 32:(kc + 1) * 32], V[:, kc * 32:(kc + 1) * 32].double()).reshape(128, -1)
            for sign, name in ((-1, "missing"), (1, "doubled")):
                d = (Pc[r0:r0 + 128] - (tru[r0:r0 + 128] + sign * part)).abs()
                print("  kc", kc, name, "fraction explained", float((d < 1e-4).float().mean()))
        break

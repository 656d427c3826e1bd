"""A/B of the headline kernel between two builds of libb200ols.so on the same box (C2 workload, device-resident):
    python tools/ab_gram.py polars_ols_b200/libb200ols.so tools/ab/libb200ols_old.so
Only ABI entry points whose signatures are identical in both builds are used."""
import ctypes as C
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from polars_ols_b200 import _lib as L  # noqa: E402  (struct layouts only)

G, n, k = 10_000, 1_000, 8
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(0)
x = torch.randn(k, G * n, dtype=torch.float64, device=dev, generator=g)
y = x.sum(0) + 0.1 * torch.randn(G * n, dtype=torch.float64, device=dev, generator=g)
offs = np.arange(G + 1, dtype=np.int64) * n
coef = torch.empty((G, k), dtype=torch.float64, device=dev)
feats = (L.Column * k)(*[L.Column(x[j].data_ptr(), None) for j in range(k)])
fr = L.Frame(n_rows=G * n, n_features=k, dtype=L.F64, memspace=L.DEVICE, add_intercept=0, target=L.Column(y.data_ptr(), None),
             features=feats, sample_weights=None, n_groups=G, group_offsets=offs.ctypes.data, row_index=None)
kw = L.OLSKwargs(1e-3, 0.0, -1, float("nan"), 0, 0, 0, 0, float("nan"))
out = L.Output(coef.data_ptr(), None)
res = {}
libs = [C.CDLL(p) for p in sys.argv[1:]]
ctxs = []
for lib in libs:
    ctx = C.c_void_p()
    assert lib.b200ols_create(0, C.byref(ctx)) == 0
    ctxs.append(ctx)
for rep in range(3):           # interleave the builds so that clocks / temperature are shared
    for path, lib, ctx in zip(sys.argv[1:], libs, ctxs):
        lib.b200ols_profile_drain.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        for _ in range(5):
            assert lib.b200ols_least_squares_coefficients(ctx, C.byref(fr), C.byref(kw), C.byref(out)) == 0
        lib.b200ols_synchronize(ctx)
        lib.b200ols_set_profiling(ctx, 1)
        for _ in range(30):
            lib.b200ols_least_squares_coefficients(ctx, C.byref(fr), C.byref(kw), C.byref(out))
        buf = np.empty(256, dtype=np.float32)
        m = lib.b200ols_profile_drain(ctx, buf.ctypes.data, 256)
        lib.b200ols_set_profiling(ctx, 0)
        res.setdefault(path, []).append(float(np.median(buf[:m])))
        ref = coef.cpu().numpy().copy()
        res.setdefault(path + ":chk", []).append(float(ref.sum()))
for p in sys.argv[1:]:
    ms = res[p]
    print(p, "gram kernel ms (3 rounds):", [round(v, 4) for v in ms], "GB/s:", round(720.64 / min(ms), 1), "checksum", res[p + ":chk"][0])

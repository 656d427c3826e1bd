"""BASELINE.json's configs at FULL size (C3, C5) / at 5M rows (C4) on the device, a sample of each checked against the
oracle, plus size-independent properties over the whole output (full-size parity belongs in `-m gpu`, not in a
tool).  C2 at full size lives in tests/test_gpu_parity.py::test_c2_full_size_normal_equations."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _gen(torch, dev, n, k, G, dtype, seed):
    g = torch.Generator(device=dev).manual_seed(seed)
    x = torch.randn(k, n, dtype=dtype, device=dev, generator=g)
    beta = 1 + 0.25 * torch.randn(G, k, dtype=torch.float64, device=dev, generator=g)
    per = n // G
    y = torch.empty(n, dtype=torch.float64, device=dev)
    step = max(1, G // 16)
    for g0 in range(0, G, step):
        g1 = min(G, g0 + step)
        xs = x[:, g0 * per:g1 * per].T.reshape(g1 - g0, per, k).to(torch.float64)
        y[g0 * per:g1 * per] = (xs * beta[g0:g1, None, :]).sum(-1).reshape(-1)
    y += 0.1 * torch.randn(n, dtype=torch.float64, device=dev, generator=g)
    return x, y.to(dtype)


def _rel(got, ref, floor=1e-3):
    got, ref = np.asarray(got, dtype=np.float64), np.asarray(ref, dtype=np.float64)
    assert (np.isnan(got) == np.isnan(ref)).all()
    m = ~np.isnan(ref)
    return float(np.max(np.abs(got[m] - ref[m]) / (floor + np.abs(ref[m]))))


def test_c3_full_size_wls_elastic_net_predictions():
    """100,000 groups x 256 rows x 16 f32 features, weights, elastic_net(alpha=1e-3, l1_ratio=0.5) predictions"""
    import torch
    import polars_ols_b200 as pls
    from polars_ols_b200 import _lib as L
    from oracle import semantics as S
    dev = torch.device("cuda", 0)
    G, per, k = 100_000, 256, 16
    x, y = _gen(torch, dev, G * per, k, G, torch.float32, 3)
    w = torch.rand(G * per, dtype=torch.float32, device=dev) + 0.05
    eng = pls.Engine(0, torch.cuda.current_stream(dev).cuda_stream or 1)
    b = pls.Batch(pls.Col(y), [pls.Col(x[i]) for i in range(k)], weights=pls.Col(w), offsets=np.arange(G + 1, dtype=np.int64) * per)
    kw = pls.OLSKwargs(alpha=1e-3, l1_ratio=0.5).to_c()
    pred = eng.least_squares(b, kw, L.PREDICTIONS, want_validity=False)[0]
    coef = eng.least_squares(b, kw, L.COEFFICIENTS)[0]
    assert bool(torch.isfinite(pred).all()) and bool(torch.isfinite(coef).all())
    # property over ALL groups: predictions == X beta with the coefficients of the coefficients-mode call.  The engine (as the
    # reference's polars expressions) rounds x * sqrt(w) to f32 before the f64 dot product, this check does not: 1e-6 relative
    chk = torch.zeros(G * per, dtype=torch.float64, device=dev)
    for j in range(k):
        chk += x[j].to(torch.float64) * coef[:, j].repeat_interleave(per)
    assert float((chk - pred).abs().max()) < 1e-6 * float(pred.abs().max() + 1)
    for g in (0, 1, G // 3, G // 2, G - 2, G - 1):                      # sample against the oracle (f32 tolerance 1e-4)
        sl = slice(g * per, (g + 1) * per)
        ref = S.least_squares(y[sl].cpu().numpy(), *[x[i, sl].cpu().numpy() for i in range(k)], sample_weights=w[sl].cpu().numpy(),
                              kwargs=S.OLSKwargs(alpha=1e-3, l1_ratio=0.5))[0]
        assert _rel(pred[sl].cpu().numpy(), ref) < 1e-4


def test_c5_full_size_lasso_coefficients():
    """1,000 groups x 10,000 rows x 64 f64 features, lasso(alpha=1e-4) coefficients"""
    import torch
    import polars_ols_b200 as pls
    from polars_ols_b200 import _lib as L
    from oracle import semantics as S
    dev = torch.device("cuda", 0)
    G, per, k = 1000, 10_000, 64
    x, y = _gen(torch, dev, G * per, k, G, torch.float64, 5)
    eng = pls.Engine(0, torch.cuda.current_stream(dev).cuda_stream or 1)
    b = pls.Batch(pls.Col(y), [pls.Col(x[i]) for i in range(k)], offsets=np.arange(G + 1, dtype=np.int64) * per)
    coef = eng.least_squares(b, pls.OLSKwargs(alpha=1e-4, l1_ratio=1.0).to_c(), L.COEFFICIENTS)[0]
    assert bool(torch.isfinite(coef).all())
    # property over ALL groups: KKT of the lasso, |x_j^T (y - X w) / n| <= alpha (+ the solver's stop tolerance) for every coordinate
    worst = 0.0
    for g0 in range(0, G, 50):
        xs = x[:, g0 * per:(g0 + 50) * per].T.reshape(50, per, k)
        r = y[g0 * per:(g0 + 50) * per].reshape(50, per) - (xs * coef[g0:g0 + 50, None, :]).sum(-1)
        grad = torch.einsum("gnk,gn->gk", xs, r) / per
        worst = max(worst, float(grad.abs().max()))
    assert worst < 1e-4 + 5e-4, worst
    for g in (0, G // 2, G - 1):
        sl = slice(g * per, (g + 1) * per)
        ref = S.solve_elastic_net(y[sl].cpu().numpy(), np.ascontiguousarray(x[:, sl].T.cpu().numpy()), 1e-4, 1.0, 1000, 1e-5, False, None)
        assert _rel(coef[g].cpu().numpy(), ref) < 1e-6


@pytest.mark.parametrize("kind", ["rolling", "rls"])
def test_c4_five_million_rows(kind):
    """one series, 5M rows x 6 f64 features: rolling_ols(252, min_periods=6) / rls(half_life=252), coefficients + predictions"""
    import torch
    import polars_ols_b200 as pls
    from polars_ols_b200 import _lib as L
    from oracle import semantics as S
    dev = torch.device("cuda", 0)
    n, k = 5_000_000, 6
    x, y = _gen(torch, dev, n, k, 1, torch.float64, 4)
    eng = pls.Engine(0, torch.cuda.current_stream(dev).cuda_stream or 1)
    b = pls.Batch(pls.Col(y), [pls.Col(x[i]) for i in range(k)])
    if kind == "rolling":
        kw = pls.RollingKwargs(window_size=252, min_periods=6, null_policy="drop").to_c()
        coef = eng.rolling_least_squares(b, kw, L.COEFFICIENTS)[0]
        pred = eng.rolling_least_squares(b, kw, L.PREDICTIONS)[0]
        okw, fn = S.RollingKwargs(window_size=252, min_periods=6, null_policy="drop"), S.rolling_least_squares
        assert bool(torch.isnan(coef[:5]).all()) and bool(torch.isfinite(coef[5:]).all())
    else:
        kw = L.RLSKwargs(252.0, 10.0, None, L.NULL_POLICY["drop"], 0)
        coef = eng.recursive_least_squares(b, kw, L.COEFFICIENTS)[0]
        pred = eng.recursive_least_squares(b, kw, L.PREDICTIONS)[0]
        okw, fn = S.RLSKwargs(half_life=252.0), S.recursive_least_squares
        assert bool(torch.isfinite(coef).all())
    # property over ALL rows: predictions == rowwise (X o Theta).sum(1) of the coefficients-mode call
    chk = sum(x[j] * coef[:, j] for j in range(k))
    ok = torch.isfinite(chk)
    assert float((chk[ok] - pred[ok]).abs().max()) < 1e-9 * float(pred[ok].abs().max() + 1)
    # head of the series (exact from row 0) and two interior samples (oracle restarted 30k rows earlier: the window / the
    # forgetting factor have forgotten the restart by then) against the sequential oracle
    xs = [x[i].cpu().numpy() for i in range(k)]
    ys = y.cpu().numpy()
    ref = fn(ys[:20_000], *[c[:20_000] for c in xs], mode="coefficients", kwargs=okw)[0]
    assert _rel(coef[:20_000].cpu().numpy(), ref) < 1e-6
    for lo in (n // 2, n - 20_000):
        sl = slice(lo - 30_000, lo + 20_000)
        ref = fn(ys[sl], *[c[sl] for c in xs], mode="coefficients", kwargs=okw)[0]
        assert _rel(coef[lo:lo + 20_000].cpu().numpy(), ref[30_000:]) < 1e-6

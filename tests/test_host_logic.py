"""CPU tests (`-m "not gpu"`): the C-ABI library loads and exports every symbol include/b200ols.h
declares, fails loudly without a device, and the host-side mirror (kwargs, grouping, sharding,
world_size-2 gather over gloo) behaves like the reference's Python layer."""
import ctypes as C
import os
import re
import socket
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent


def test_library_exports_every_declared_symbol():
    from polars_ols_b200 import _lib
    header = (ROOT / "include" / "b200ols.h").read_text()
    declared = set(re.findall(r"B200OLS_API[^;(]*?\b(b200ols_[a-z_0-9]+)\s*\(", header))
    assert declared == set(_lib.EXPORTS), declared ^ set(_lib.EXPORTS)
    L = C.CDLL(str(_lib.SO_PATH))
    for name in declared:
        assert hasattr(L, name), name
    assert _lib.load().b200ols_version() == 100


def test_struct_layouts_match_the_header():
    from polars_ols_b200 import _lib
    assert C.sizeof(_lib.Column) == 16
    assert C.sizeof(_lib.Frame) == 8 + 16 + 16 + 8 + 8 + 8 + 8 + 8 + 8
    assert _lib.Frame.row_index_on_device.offset == 80
    assert C.sizeof(_lib.KeyColumn) == 16 and C.sizeof(_lib.GroupPlan) == 48
    assert C.sizeof(_lib.OLSKwargs) == 56 and C.sizeof(_lib.RLSKwargs) == 40 and C.sizeof(_lib.RollingKwargs) == 32
    assert _lib.Frame.target.offset == 24 and _lib.Frame.group_offsets.offset == 64


def test_no_cpu_fallback_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    import polars_ols_b200 as pls
    with pytest.raises(pls.B200OLSError) as ei:
        pls.Engine(0)
    assert ei.value.code == -4 and "no CPU fallback" in str(ei.value)
    d = {"y": np.zeros(4), "x": np.ones(4)}
    with pytest.raises(pls.B200OLSError):
        pls.Frame(d).select(pls.col("y").least_squares.ols("x"))


def test_product_never_imports_the_oracle():
    """the oracle is test infrastructure: nothing under polars_ols_b200/ may import, load or link it"""
    # code references only (comments may cite the checker): python imports, C includes, dlopen names
    bad = re.compile(r"^\s*(import\s+oracle|from\s+oracle|from\s+\.\.?oracle)|#\s*include\s*\"[^\"]*(oracle|hostcheck)|libols_oracle|libhostcheck", re.M)
    for p in (ROOT / "polars_ols_b200").rglob("*"):
        if p.suffix in (".py", ".cu", ".cuh", ".h"):
            assert not bad.search(p.read_text()), p
    # and the shared library does not link against it
    import subprocess
    out = subprocess.run(["ldd", str(ROOT / "polars_ols_b200" / "libb200ols.so")], capture_output=True, text=True).stdout
    assert "oracle" not in out


def test_kwargs_mirror_reference_defaults_and_validation():
    import polars_ols_b200 as pls
    k = pls.OLSKwargs()
    assert (k.alpha, k.l1_ratio, k.max_iter, k.tol, k.positive, k.solve_method, k.rcond, k.null_policy) == \
        (0.0, None, 1000, 1e-5, False, None, None, "ignore")
    assert pls.RLSKwargs().null_policy == "drop" and pls.RLSKwargs().initial_state_covariance == 10.0
    assert pls.RollingKwargs().null_policy == "drop_window" and pls.RollingKwargs().window_size == 1_000_000
    with pytest.raises(AssertionError):
        pls.OLSKwargs(null_policy="drop_window")          # polars_ols/least_squares.py:111-115
    with pytest.raises(AssertionError):
        pls.OLSKwargs(solve_method="nope")
    c = pls.OLSKwargs(alpha=None, l1_ratio=None, max_iter=None, tol=None).to_c()
    assert np.isnan(c.alpha) and np.isnan(c.l1_ratio) and c.max_iter == -1 and np.isnan(c.tol)
    ns = pls.col("y").least_squares
    assert ns.ridge("x", alpha=1.0).kwargs.l1_ratio == 0.0      # polars_ols/__init__.py:123
    assert ns.lasso("x", alpha=1.0).kwargs.l1_ratio == 1.0      # :132
    assert ns.elastic_net("x", alpha=1.0).kwargs.l1_ratio == 0.5
    assert ns.rls("x").kwargs.null_policy == "drop" and ns.rolling_ols("x", window_size=5).kwargs.null_policy == "drop"
    assert ns.expanding_ols("x").kwargs.half_life is None


def test_group_plan_matches_polars_over_semantics():
    from polars_ols_b200.least_squares import _group_plan
    g = np.array([2, 0, 2, 1, 0, 2])
    keys, offs, ridx, inv = _group_plan([g])
    assert keys.tolist() == [0, 1, 2] and offs.tolist() == [0, 2, 3, 6]
    assert ridx.tolist() == [1, 4, 3, 0, 2, 5]                  # stable: frame order inside each group
    keys, offs, ridx, inv = _group_plan([np.array([0, 0, 1, 1, 1])])
    assert ridx is None and offs.tolist() == [0, 2, 5]          # contiguous slices need no gather


def test_columns_from_arrow_and_masks():
    import pyarrow as pa
    from polars_ols_b200 import as_col
    c = as_col(pa.array([1.0, None, 3.0, 4.0, None, 6.0, 7.0, 8.0, 9.0]))
    assert c.values.dtype == np.float64 and c.validity is not None
    assert np.unpackbits(c.validity, bitorder="little")[:9].tolist() == [1, 0, 1, 1, 0, 1, 1, 1, 1]
    c = as_col((np.arange(4.0), np.array([True, True, True, True])))
    assert c.validity is None
    c = as_col(np.arange(5))
    assert c.values.dtype == np.float64


def test_shard_groups_balances_rows():
    from polars_ols_b200.parallel import shard_groups
    offs = np.concatenate([[0], np.cumsum([10, 1000, 10, 10, 1000, 970])])
    sh = shard_groups(offs, 2)
    assert sh[0][0] == 0 and sh[-1][1] == 6 and sh[0][1] == sh[1][0]
    rows = [offs[b] - offs[a] for a, b in sh]
    assert abs(rows[0] - rows[1]) <= 1000
    assert shard_groups(np.array([0, 5]), 4)[-1] == (1, 1) or sum(b - a for a, b in shard_groups(np.array([0, 5]), 4)) == 1


def _gloo_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    from polars_ols_b200.parallel import gather_group_results, shard_groups
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    offs = np.concatenate([[0], np.cumsum([5, 50, 5, 40, 7])])
    shards = shard_groups(offs, world)
    a, b = shards[rank]
    local = torch.arange(a, b, dtype=torch.float64)[:, None] * torch.ones(1, 3, dtype=torch.float64)  # "coefficients" = group id
    full = gather_group_results(local, shards)
    # equal shards take the single-collective path (one all_gather_into_tensor straight into `out`)
    eq = [(0, 3), (3, 6)]
    loc = torch.arange(eq[rank][0], eq[rank][1], dtype=torch.float64)[:, None] * torch.ones(1, 2, dtype=torch.float64)
    out = torch.empty((6, 2), dtype=torch.float64)
    full_eq = gather_group_results(loc, eq, out=out)
    assert full_eq.data_ptr() == out.data_ptr() and full_eq[:, 0].tolist() == [0.0, 1.0, 2.0, 3.0, 4.0, 5.0]
    # the one exchange of a time-sharded rls: every rank's k*k + k + 1 state map, one tensor all-gather
    from polars_ols_b200.parallel import _all_gather_array
    maps = _all_gather_array(np.arange(7, dtype=np.float64) + 10 * rank)
    assert len(maps) == world and all(np.array_equal(maps[r], np.arange(7) + 10.0 * r) for r in range(world))
    q.put((rank, full[:, 0].tolist()))
    dist.destroy_process_group()


def test_world_size_2_gather_over_gloo():
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in ps]
    res = [q.get(timeout=120) for _ in ps]
    [p.join(60) for p in ps]
    for rank, vals in res:
        assert vals == [0.0, 1.0, 2.0, 3.0, 4.0]


# ---- time-axis sharding of one long series -------------------------------------------
def _rolling_case(rng, n, k, null_frac, lead_nulls):
    x = rng.standard_normal((n, k))
    y = x @ (1 + 0.25 * rng.standard_normal(k)) + 0.1 * rng.standard_normal(n)
    valid = rng.random(n) >= null_frac
    valid[:lead_nulls] = False
    return y, x, valid


@pytest.mark.parametrize("null_policy", ["drop", "drop_window"])
@pytest.mark.parametrize("null_frac,lead_nulls", [(0.0, 0), (0.3, 0), (0.7, 0), (0.2, 37), (0.95, 0)])
def test_rolling_halo_reproduces_whole_series(null_policy, null_frac, lead_nulls):
    """rows >= start computed from [halo_start, n) equal the whole-series result (oracle on both sides)."""
    from oracle import semantics as S
    from polars_ols_b200.parallel import rolling_halo_start

    rng = np.random.default_rng(7)
    n, k = 600, 3
    fixed = null_policy == "drop_window"
    for trial in range(12):
        W = int(rng.integers(k + 5, 40))            # min_periods > k keeps every solved window well-posed, so the
        mp = int(rng.integers(k + 2, W + 1))        # comparison is not between two roundings of a singular solve
        y, x, valid = _rolling_case(rng, n, k, null_frac, lead_nulls)
        yz, xz = np.where(valid, y, 0.0), np.where(valid[:, None], x, 0.0)
        full = S.solve_rolling_ols(yz, xz, W, mp, None, None, valid, null_policy)
        vf = None if valid.all() else (lambda a, b: valid[a:b].copy())
        for start in (8, 64, 200, 424, 592):
            h = rolling_halo_start(vf, start, W, mp, fixed, align=8)
            assert 0 <= h <= start and h % 8 == 0
            part = S.solve_rolling_ols(np.ascontiguousarray(yz[h:]), np.ascontiguousarray(xz[h:]), W, mp, None, None,
                                       np.ascontiguousarray(valid[h:]), null_policy)
            np.testing.assert_allclose(part[start - h:], full[start:], rtol=1e-8, atol=1e-10, equal_nan=True)
        if null_frac == 0.0:  # without nulls a bounded halo always suffices
            assert rolling_halo_start(vf, 592, W, mp, fixed, align=8) > 0 or 592 <= 2 * W + 16 + mp


def test_rls_shard_maps_compose_to_the_sequential_filter():
    """information-form maps of consecutive shards, folded over the prior, give the state the sequential
    reference filter reaches (src/least_squares.rs:505-540)."""
    from oracle import semantics as S
    from polars_ols_b200.parallel import rls_entering_state, rls_prior_information, shard_rows

    rng = np.random.default_rng(3)
    n, k = 500, 4
    x = rng.standard_normal((n, k))
    y = x @ np.arange(1, k + 1) + 0.1 * rng.standard_normal(n)
    valid = rng.random(n) > 0.1
    for half_life, mean in ((None, None), (30.0, [0.5] * k)):
        lam = 1.0 if half_life is None else np.exp(np.log(0.5) / half_life)
        coef = S.solve_recursive_least_squares(np.where(valid, y, 0), np.where(valid[:, None], x, 0), half_life, 10.0, mean, valid)
        shards = shard_rows(n, 4, align=8)
        maps = []
        for a, b in shards:
            A, bb, D = np.zeros((k, k)), np.zeros(k), 1.0
            for r in range(a, b):
                if valid[r]:
                    A, bb, D = lam * A + np.outer(x[r], x[r]), lam * bb + x[r] * y[r], D * lam
            maps.append(np.concatenate([A.reshape(-1), bb, [D]]))
        maps = np.stack(maps)
        prior = rls_prior_information(k, 10.0, mean)
        for rank in range(1, 4):
            st = rls_entering_state(maps, prior, rank)
            theta = np.linalg.solve(st[:k * k].reshape(k, k), st[k * k:])
            np.testing.assert_allclose(theta, coef[shards[rank][0] - 1], rtol=1e-8, atol=1e-10)


def test_shard_rows_alignment_and_cover():
    from polars_ols_b200.parallel import shard_rows
    for n in (0, 1, 63, 64, 1000, 50_000_000):
        for w in (1, 2, 3, 8):
            sh = shard_rows(n, w)
            assert sh[0][0] == 0 and sh[-1][1] == n
            assert all(a % 64 == 0 for a, _ in sh) and all(sh[i][1] == sh[i + 1][0] for i in range(w - 1))


def test_formula_parser_subset():
    """polars_ols/utils.py:62-111 (patsy subset): columns, ':' interactions, '-1' switches the intercept off."""
    import polars_ols_b200 as pls
    ex, icpt = pls.build_expressions_from_patsy_formula("y ~ x1 + x2 + x3:x4", include_dependent_variable=True)
    assert [e.output_name for e in ex] == ["y", "x1", "x2", "x3:x4"] and icpt
    ex, icpt = pls.build_expressions_from_patsy_formula("x1 + x2 -1")
    assert [e.output_name for e in ex] == ["x1", "x2"] and not icpt
    f = pls.Frame({"x3": np.arange(4.0), "x4": (np.arange(4.0) + 1, np.array([True, False, True, True]))})
    c = pls.build_expressions_from_patsy_formula("x3:x4")[0][0].resolve(f)
    np.testing.assert_array_equal(c.values, [0.0, 2.0, 6.0, 12.0])
    assert np.unpackbits(c.validity, bitorder="little")[:4].tolist() == [1, 0, 1, 1]    # null in a factor -> null
    for bad in ("log(x1)", "C(group) + x1", "x1 * x2"):
        with pytest.raises(NotImplementedError):
            pls.build_expressions_from_patsy_formula(bad)
    with pytest.raises(AssertionError):
        pls.build_expressions_from_patsy_formula("y ~ x1")                              # LHS not allowed here
    with pytest.raises(AssertionError):
        pls.build_expressions_from_patsy_formula("x1 + x2", include_dependent_variable=True)
    # the reference tests `"-1" not in formula` on the RAW text (polars_ols/utils.py:99): "- 1" with a space KEEPS the
    # intercept column although patsy dropped the intercept term — copied, not fixed; repeated terms appear once
    e = pls.col("y").least_squares.from_formula("x1 + x2 - 1", window_size=20)
    assert e.kind == "rolling_least_squares" and e.add_intercept
    e = pls.col("y").least_squares.from_formula("x1 + x2 -1", window_size=20)
    assert not e.add_intercept
    ex, _ = pls.build_expressions_from_patsy_formula("x1 + x2 + x1 + x2:x3 + x3:x2")
    assert [t.output_name for t in ex] == ["x1", "x2", "x2:x3"]
    e = pls.compute_least_squares_from_formula("y ~ x1", half_life=3.0)
    assert e.kind == "recursive_least_squares" and e.add_intercept


def test_cd_stop_threshold_is_equivalent_to_the_sqrt_test():
    """cd_solve.cuh replaces the reference's stop test `norm_l2(w - w_old) < tol` (src/least_squares.rs:436-444), i.e.
    sqrt(d2) < tol, by d2 <= cut with cut = the largest double whose correctly rounded square root is below tol.
    Restated here bit for bit (numpy's float64 sqrt is correctly rounded, like the device's): the two tests must agree
    for every d2 around the boundary."""
    def cut_of(tol):
        cut = np.float64(tol) * np.float64(tol)
        bits = lambda v: np.float64(v).view(np.int64)                      # noqa: E731
        val = lambda b: np.int64(b).view(np.float64)                       # noqa: E731
        while cut > 0.0 and np.sqrt(cut) >= tol:
            cut = val(bits(cut) - 1)
        while cut < 1.0e300 and np.sqrt(val(bits(cut) + 1)) < tol:
            cut = val(bits(cut) + 1)
        return cut

    rng = np.random.default_rng(0)
    tols = [1e-5, 1e-4, 1e-8, 1e-3, 0.1, 1.0, 3.0, 1e-12] + list(10.0 ** rng.uniform(-12, 2, size=200))
    for tol in tols:
        tol = np.float64(tol)
        cut = cut_of(tol)
        assert np.sqrt(cut) < tol
        b = cut.view(np.int64)
        for off in range(-4, 5):
            d2 = np.int64(b + off).view(np.float64)
            assert (np.sqrt(d2) < tol) == (d2 <= cut), (tol, off)
        for d2 in (0.0, tol * tol * 0.5, tol * tol * 2.0, np.inf):
            assert (np.sqrt(np.float64(d2)) < tol) == (np.float64(d2) <= cut)
        assert not (np.float64(np.nan) <= cut)                                 # NaN never stops the sweeps, as sqrt(NaN) < tol


def test_expressions_are_immutable_like_polars():
    """`.over()` / `.alias()` return new expressions (polars semantics): re-using `e` after `e.over(..)` stays ungrouped."""
    import polars_ols_b200 as pls
    e = pls.col("y").least_squares.ols("x1", "x2")
    g = e.over("group")
    a = g.alias("pred")
    assert e._over == [] and g._over == ["group"] and a._over == ["group"]
    assert e.output_name == "y" and g.output_name == "y" and a.output_name == "pred"
    p = pls.col("coefficients").least_squares.predict("x1", "x2", name="p0")
    q = p.alias("p1")
    assert p.output_name == "p0" and q.output_name == "p1"


def test_cd_branch_free_soft_threshold_is_the_reference_formula():
    """cd_solve.cuh evaluates soft_threshold (src/least_squares.rs:373-379: signum(x) * max(|x| - t, 0), then max(., 0)
    when `positive`) as: keep = |x| - t > 0 (and x > 0 when positive); value = copysign(|x| - t, x) if keep else 0.
    Restated with numpy: the same numbers for every input class (zeros compare equal regardless of sign; a NaN gives 0
    in both, as f64::max ignores it)."""
    rng = np.random.default_rng(1)
    x = np.concatenate([rng.normal(size=2000) * 10.0 ** rng.integers(-8, 4, size=2000), [0.0, -0.0, np.nan, np.inf, -np.inf, 1e-320, -1e-320]])
    for t in (0.0, 1e-12, 1e-4, 0.5, 3.0, 1e6):
        for positive in (False, True):
            with np.errstate(invalid="ignore"):
                ref = np.copysign(np.fmax(np.abs(x) - t, 0.0), x)          # f64::max == fmax: NaN is ignored
                if positive:
                    ref = np.fmax(ref, 0.0)
                av = np.abs(x) - t
                keep = av > 0.0
                if positive:
                    keep = keep & (x > 0.0)
                new = np.where(keep, np.copysign(av, x), 0.0)
            assert np.array_equal(ref + 0.0, new + 0.0), (t, positive)     # + 0.0: -0.0 and 0.0 are the same coefficient


def test_polars_adapter_is_guarded():
    """the batched `.over()` route for real polars frames (polars_ols/least_squares.py:199-239 replaced by ONE engine call)
    imports without polars and says so"""
    from polars_ols_b200 import polars_adapter as pa
    import polars_ols_b200 as pls
    e = pls.col("y").least_squares.ridge("x1", "x2:x3", alpha=1.0, sample_weights="w", mode="coefficients").over("g", "h")
    assert pa._needed_columns(e) == ["y", "x1", "x2:x3", "w", "g", "h"] or pa._needed_columns(e)[:2] == ["y", "x1"]
    if not pa.available():
        with pytest.raises(ImportError):
            pa.over_batched(None, e)
        assert pa.register_namespace() is None

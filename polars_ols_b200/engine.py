"""Python handle on a ``b200ols_ctx`` + marshalling of columns into the C-ABI frame descriptor.

Mirrors what ``src/expressions.rs`` does between polars Series and the solvers (cast, validity,
grouping) — but only describes the buffers; no arithmetic happens on the host.
"""
from __future__ import annotations

import ctypes as C
import math
from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple

import numpy as np

from . import _lib as L

try:  # torch is plumbing (device tensors / streams), not required for host frames
    import torch
except Exception:  # pragma: no cover
    torch = None


def _is_torch(x) -> bool:
    return torch is not None and isinstance(x, torch.Tensor)


@dataclass
class Col:
    """One input column: values buffer (numpy array or CUDA torch tensor) + optional Arrow validity bitmap."""
    values: object
    validity: object = None  # np.uint8 bitmap (host) / torch.uint8 tensor (device) / None

    @property
    def is_device(self) -> bool:
        return _is_torch(self.values) and self.values.is_cuda

    def __len__(self):
        return int(self.values.shape[0])


def _bitmap_from_bool(mask: np.ndarray) -> Optional[np.ndarray]:
    mask = np.asarray(mask, dtype=bool)
    if mask.all():
        return None
    return np.packbits(mask, bitorder="little")


def as_col(x) -> Col:
    """numpy array | (values, bool mask) | numpy masked array | pyarrow Array | torch tensor -> Col."""
    if isinstance(x, Col):
        return x
    if _is_torch(x):
        if x.is_cuda:
            t = x if x.dtype in (torch.float32, torch.float64) else x.to(torch.float64)
            return Col(t.contiguous())
        x = x.numpy()
    if isinstance(x, tuple) and len(x) == 2:
        v, m = x
        c = as_col(v)
        if m is not None:
            c = Col(c.values, _bitmap_from_bool(m))
        return c
    if isinstance(x, np.ma.MaskedArray):
        return Col(_float_array(np.ma.getdata(x)), _bitmap_from_bool(~np.ma.getmaskarray(x)))
    try:
        import pyarrow as pa
        if isinstance(x, pa.ChunkedArray):
            x = x.combine_chunks()
        if isinstance(x, pa.Array):
            if not (pa.types.is_floating(x.type) and x.type.bit_width in (32, 64)):
                x = x.cast(pa.float64())
            bufs = x.buffers()
            dt = np.float32 if x.type.bit_width == 32 else np.float64
            vals = np.frombuffer(bufs[1], dtype=dt, count=len(x) + x.offset)[x.offset:]
            validity = None
            if x.null_count > 0:
                if x.offset == 0:
                    validity = np.frombuffer(bufs[0], dtype=np.uint8, count=(len(x) + 7) // 8)
                else:
                    validity = _bitmap_from_bool(np.asarray(x.is_valid()))
            return Col(vals, validity)
    except ImportError:  # pragma: no cover
        pass
    return Col(_float_array(np.asarray(x)))


def _float_array(a: np.ndarray) -> np.ndarray:
    if a.dtype not in (np.float32, np.float64):
        a = a.astype(np.float64)  # cast(&DataType::Float64), src/expressions.rs:33,47,80
    return np.ascontiguousarray(a)


@dataclass
class DevicePtr:
    """A raw device pointer owned by the engine (e.g. the permutation b200ols_group_plan_build leaves on the device)."""
    ptr: int


@dataclass
class DevicePlan:
    """Result of ``Engine.group_plan``: CSR groups of an `.over()` key set, planned on the device
    (``b200ols_group_plan_build``).  ``offsets`` / ``first_row`` are host arrays; the permutation stays on the device
    (``row_index``: DevicePtr, or None when the groups are contiguous row slices) and is valid until the engine's next plan."""
    engine: "Engine"
    n_rows: int
    n_groups: int
    offsets: np.ndarray
    first_row: np.ndarray
    row_index: Optional[DevicePtr]
    device_ms: float
    serial: int = 0

    def _check(self):
        if self.engine._plan_serial != self.serial:
            raise RuntimeError("this group plan was superseded by a later plan on the same engine")

    def row_index_numpy(self) -> np.ndarray:
        self._check()
        out = np.empty(self.n_rows, dtype=np.int64)
        L.check(self.engine._lib.b200ols_group_plan_row_index(self.engine._ctx, out.ctypes.data, L.HOST))
        return out

    def group_of_row(self) -> np.ndarray:
        """group id of every original row (int32) — the broadcast map of `.over()`"""
        self._check()
        out = np.empty(self.n_rows, dtype=np.int32)
        L.check(self.engine._lib.b200ols_group_plan_group_of_row(self.engine._ctx, out.ctypes.data, L.HOST))
        return out


_KEY_DTYPES = {"int64": L.KEY_I64, "int32": L.KEY_I32, "uint64": L.KEY_U64, "uint32": L.KEY_U32,
               "float64": L.KEY_F64, "float32": L.KEY_F32}


def _key_column(k):
    """key column -> (buffer kept alive, pointer, B200OLS_KEY_* dtype, is_device).  Small integer / bool keys are widened
    to int32; anything else (strings, objects, datetimes) is dictionary-encoded on the host first — the codes ascend
    with the key, so the device plan orders the groups as numpy.unique would."""
    if _is_torch(k):
        if k.is_cuda:
            name = str(k.dtype).replace("torch.", "")
            if name not in _KEY_DTYPES:
                k = k.to(torch.int32 if name in ("int8", "int16", "uint8", "bool") else torch.int64)
                name = str(k.dtype).replace("torch.", "")
            k = k.contiguous()
            return k, k.data_ptr(), _KEY_DTYPES[name], True
        k = k.numpy()
    k = np.asarray(k)
    if k.dtype.name not in _KEY_DTYPES:
        if k.dtype.kind in "iub" and k.dtype.itemsize < 4:
            k = k.astype(np.int32)
        else:
            _, k = np.unique(k, return_inverse=True)
            k = k.reshape(-1).astype(np.int64)
    k = np.ascontiguousarray(k)
    return k, k.ctypes.data, _KEY_DTYPES[k.dtype.name], False


@dataclass
class Batch:
    """All inputs of one `.over()` evaluation, as the C ABI wants them."""
    target: Col
    features: List[Col]
    weights: Optional[Col] = None
    add_intercept: bool = False
    offsets: Optional[np.ndarray] = None    # int64 [G+1], host
    row_index: object = None                # int64 [N] (numpy or CUDA tensor), a DevicePtr (Engine.group_plan), or None
    n_groups: int = 1

    def harmonise(self) -> Tuple[int, int]:
        """common dtype (polars supertype) + common memory space; returns (dtype, memspace)."""
        cols = [self.target, *self.features] + ([self.weights] if self.weights is not None else [])
        dev = [c.is_device for c in cols]
        if any(dev) and not all(dev):
            raise ValueError("all columns of a frame must live in the same memory space (host or CUDA)")
        n = len(self.target)
        for c in cols:
            if len(c) != n:
                raise ValueError("all input series passed must be of equal length")  # src/expressions.rs:96-100
        if all(dev):
            f32 = all(c.values.dtype == torch.float32 for c in cols)
            want = torch.float32 if f32 else torch.float64
            for c in cols:
                if c.values.dtype != want:
                    c.values = c.values.to(want)
            return (L.F32 if f32 else L.F64), L.DEVICE
        f32 = all(c.values.dtype == np.float32 for c in cols)
        want = np.float32 if f32 else np.float64
        for c in cols:
            if c.values.dtype != want or not c.values.flags.c_contiguous:
                c.values = np.ascontiguousarray(c.values, dtype=want)
        return (L.F32 if f32 else L.F64), L.HOST


def _ptr(x) -> Optional[int]:
    if x is None:
        return None
    if _is_torch(x):
        return x.data_ptr()
    return x.ctypes.data


class Engine:
    """One ``b200ols_ctx``: a device, a stream, and the engine's device scratch."""

    def __init__(self, device: int = 0, stream: Optional[int] = None):
        self._lib = L.load()
        self._ctx = C.c_void_p()
        if stream is None:
            L.check(self._lib.b200ols_create(device, C.byref(self._ctx)))
        else:
            L.check(self._lib.b200ols_create_on_stream(device, C.c_void_p(stream), C.byref(self._ctx)))
        self.device = device
        self._plan_serial = 0

    def group_plan(self, keys: Sequence) -> DevicePlan:
        """``b200ols_group_plan_build``: `.over()` key column(s) (numpy arrays or CUDA tensors) -> CSR groups, planned on
        the device.  Groups ascend by key tuple, rows keep the frame's order inside a group."""
        cols = [_key_column(k) for k in keys]
        dev = [c_[3] for c_ in cols]
        if any(dev) and not all(dev):
            raise ValueError("all key columns must live in the same memory space (host or CUDA)")
        n = int(cols[0][0].shape[0])
        for c_ in cols:
            if int(c_[0].shape[0]) != n:
                raise ValueError("all key columns must be of equal length")
        kc = (L.KeyColumn * len(cols))(*[L.KeyColumn(c_[1], c_[2], 0) for c_ in cols])
        gp = L.GroupPlan()
        L.check(self._lib.b200ols_group_plan_build(self._ctx, kc, len(cols), n, L.DEVICE if all(dev) else L.HOST, C.byref(gp)))
        self._plan_serial += 1
        G = int(gp.n_groups)
        offsets = np.ctypeslib.as_array(C.cast(gp.group_offsets, C.POINTER(C.c_int64)), shape=(G + 1,)).copy()
        first = (np.ctypeslib.as_array(C.cast(gp.group_first_row, C.POINTER(C.c_int64)), shape=(G,)).copy()
                 if G > 0 else np.empty(0, dtype=np.int64))
        del cols
        return DevicePlan(self, n, G, offsets, first, DevicePtr(gp.row_index) if gp.row_index else None, float(gp.device_ms),
                          self._plan_serial)

    def close(self):
        if self._ctx:
            self._lib.b200ols_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass

    # -- misc ---------------------------------------------------------------------------------------
    def synchronize(self):
        L.check(self._lib.b200ols_synchronize(self._ctx))

    @property
    def launch_count(self) -> int:
        return int(self._lib.b200ols_launch_count(self._ctx))

    def set_tuning(self, tile_rows: int = 0, warps_per_cta: int = 0, ctas_per_sm: int = 0):
        L.check(self._lib.b200ols_set_tuning(self._ctx, tile_rows, warps_per_cta, ctas_per_sm))

    def set_variant(self, variant: int = 0, unroll: int = 0):
        L.check(self._lib.b200ols_set_variant(self._ctx, variant, unroll))

    def set_profiling(self, enabled: bool):
        L.check(self._lib.b200ols_set_profiling(self._ctx, 1 if enabled else 0))

    def profile_drain(self, max_n: int = 4096) -> np.ndarray:
        """device durations (ms) of the dominant kernel of every call since the last drain"""
        buf = np.empty(max_n, dtype=np.float32)
        n = self._lib.b200ols_profile_drain(self._ctx, buf.ctypes.data, max_n)
        if n < 0:
            L.check(n)
        return buf[:n].copy()

    def last_group_flags(self, n_groups: int) -> np.ndarray:
        out = np.empty(n_groups, dtype=np.int32)
        L.check(self._lib.b200ols_last_group_flags(self._ctx, out.ctypes.data, n_groups))
        return out

    def pinned_empty(self, shape, dtype=np.float64) -> np.ndarray:
        """numpy array over page-locked host memory (freed with the process); for per-call uploads."""
        n = int(np.prod(shape)) * np.dtype(dtype).itemsize
        p = self._lib.b200ols_host_alloc(max(n, 1))
        if not p:
            raise L.B200OLSError(L.ERR_CUDA, self._lib.b200ols_last_error().decode())
        buf = (C.c_char * max(n, 1)).from_address(p)
        return np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)

    # -- frame marshalling ----------------------------------------------------------------------------
    def _frame(self, b: Batch):
        dtype, memspace = b.harmonise()
        n = len(b.target)
        keep = []
        feats = (L.Column * max(len(b.features), 1))()
        for j, c in enumerate(b.features):
            feats[j].values = _ptr(c.values)
            feats[j].validity = _ptr(c.validity)
        fr = L.Frame()
        fr.n_rows = n
        fr.n_features = len(b.features)
        fr.dtype = dtype
        fr.memspace = memspace
        fr.add_intercept = 1 if b.add_intercept else 0
        fr.target.values = _ptr(b.target.values)
        fr.target.validity = _ptr(b.target.validity)
        fr.features = feats
        wcol = None
        if b.weights is not None:
            wcol = L.Column(_ptr(b.weights.values), _ptr(b.weights.validity))
            fr.sample_weights = C.pointer(wcol)
        offs = None
        if b.offsets is not None:
            offs = np.ascontiguousarray(b.offsets, dtype=np.int64)
            fr.n_groups = len(offs) - 1
            fr.group_offsets = offs.ctypes.data
        else:
            fr.n_groups = 1
        ridx = b.row_index
        if isinstance(ridx, DevicePtr):
            fr.row_index = ridx.ptr
            fr.row_index_on_device = 1
        elif ridx is not None:
            if memspace == L.DEVICE:
                if not _is_torch(ridx):
                    ridx = torch.as_tensor(np.ascontiguousarray(ridx, dtype=np.int64), device=b.target.values.device)
                ridx = ridx.to(torch.int64).contiguous()
            else:
                ridx = np.ascontiguousarray(ridx, dtype=np.int64)
            fr.row_index = _ptr(ridx)
        keep += [feats, wcol, offs, ridx, b]
        return fr, keep, dtype, memspace, n, int(fr.n_groups)

    def _alloc_out(self, memspace, shape, want_validity, like):
        if memspace == L.DEVICE:
            v = torch.empty(shape, dtype=torch.float64, device=like.device)
            m = torch.empty(shape, dtype=torch.uint8, device=like.device) if want_validity else None
        else:
            v = np.empty(shape, dtype=np.float64)
            m = np.empty(shape, dtype=np.uint8) if want_validity else None
        out = L.Output(_ptr(v), _ptr(m))
        return out, v, m

    # -- the six entry points ---------------------------------------------------------------------------
    def least_squares(self, b: Batch, kw: L.OLSKwargs, mode: int, want_validity: bool = True, out=None):
        fr, keep, dtype, memspace, n, G = self._frame(b)
        F = len(b.features) + (1 if b.add_intercept else 0)
        shape = (G, F) if mode == L.COEFFICIENTS else (n,)
        if out is None:
            o, v, m = self._alloc_out(memspace, shape, want_validity and mode != L.COEFFICIENTS, b.target.values)
        else:
            v, m = out
            o = L.Output(_ptr(v), _ptr(m))
        if mode == L.COEFFICIENTS:
            L.check(self._lib.b200ols_least_squares_coefficients(self._ctx, C.byref(fr), C.byref(kw), C.byref(o)))
        else:
            L.check(self._lib.b200ols_least_squares(self._ctx, C.byref(fr), C.byref(kw), mode, C.byref(o)))
        del keep
        return v, m

    def least_squares_statistics(self, b: Batch, kw: L.OLSKwargs) -> dict:
        """b200ols_least_squares_statistics: per-group r2 / mae / mse [G] and coefficients / standard_errors / t_values /
        p_values [G, F] (numpy for host frames, CUDA tensors for device frames)."""
        fr, keep, dtype, memspace, n, G = self._frame(b)
        F = len(b.features) + (1 if b.add_intercept else 0)
        names = ("r2", "mae", "mse", "coefficients", "standard_errors", "t_values", "p_values")
        arrs = {}
        for nm in names:
            shape = (G,) if nm in ("r2", "mae", "mse") else (G, F)
            arrs[nm] = (torch.empty(shape, dtype=torch.float64, device=b.target.values.device) if memspace == L.DEVICE
                        else np.empty(shape, dtype=np.float64))
        so = L.StatisticsOutput(*[_ptr(arrs[nm]) for nm in names])
        L.check(self._lib.b200ols_least_squares_statistics(self._ctx, C.byref(fr), C.byref(kw), C.byref(so)))
        del keep
        return arrs

    def multi_target_least_squares(self, b: Batch, targets: List[Col], kw: L.OLSKwargs, mode: int):
        """b200ols_multi_target_least_squares: values / validity of shape [n_targets, n_rows] (target-major)."""
        tcols = list(targets)
        b2 = Batch(tcols[-1], list(b.features) + tcols[:-1], b.weights, b.add_intercept, b.offsets, b.row_index, b.n_groups)
        dtype, memspace = b2.harmonise()      # one dtype / memory space over features, targets and weights
        b.target = tcols[-1]
        fr, keep, dtype, memspace, n, G = self._frame(b)
        m = len(tcols)
        tt = (L.Column * m)(*[L.Column(_ptr(t.values), _ptr(t.validity)) for t in tcols])
        o, v, msk = self._alloc_out(memspace, (m, n), True, tcols[0].values)
        L.check(self._lib.b200ols_multi_target_least_squares(self._ctx, C.byref(fr), m, tt, C.byref(kw), mode, C.byref(o)))
        del keep, b2
        return v, msk

    def prepare_least_squares(self, b: Batch, kw: L.OLSKwargs, mode: int, out_values, out_validity=None):
        """Marshal once, call many times: returns a zero-argument callable that re-issues the same C-ABI
        call (same buffers).  Used by steady-state loops (bench.py) to keep Python out of the timed region."""
        fr, keep, dtype, memspace, n, G = self._frame(b)
        o = L.Output(_ptr(out_values), _ptr(out_validity))
        fn = self._lib.b200ols_least_squares_coefficients if mode == L.COEFFICIENTS else None
        ctx, lib = self._ctx, self._lib
        frp, kwp, op = C.byref(fr), C.byref(kw), C.byref(o)

        def call(_keep=(keep, fr, kw, o, out_values, out_validity)):
            rc = fn(ctx, frp, kwp, op) if fn is not None else lib.b200ols_least_squares(ctx, frp, kwp, mode, op)
            if rc != 0:
                L.check(rc)

        return call

    def recursive_least_squares(self, b: Batch, kw: L.RLSKwargs, mode: int, mean_keepalive=None):
        fr, keep, dtype, memspace, n, G = self._frame(b)
        F = len(b.features) + (1 if b.add_intercept else 0)
        shape = (n, F) if mode == L.COEFFICIENTS else (n,)
        o, v, m = self._alloc_out(memspace, shape, True, b.target.values)
        if mode == L.COEFFICIENTS:
            L.check(self._lib.b200ols_recursive_least_squares_coefficients(self._ctx, C.byref(fr), C.byref(kw), C.byref(o)))
        else:
            L.check(self._lib.b200ols_recursive_least_squares(self._ctx, C.byref(fr), C.byref(kw), mode, C.byref(o)))
        del keep, mean_keepalive
        return v, m

    def recursive_least_squares_state(self, b: Batch, kw: L.RLSKwargs, keepalive=None) -> np.ndarray:
        """b200ols_recursive_least_squares_state: [G, F*F + F + 1] information state leaving each series
        (A row-major symmetric, b, decay D) given kw.initial_information (or the prior)."""
        fr, keep, dtype, memspace, n, G = self._frame(b)
        F = len(b.features) + (1 if b.add_intercept else 0)
        out = np.empty((G, F * F + F + 1), dtype=np.float64)
        L.check(self._lib.b200ols_recursive_least_squares_state(self._ctx, C.byref(fr), C.byref(kw), out.ctypes.data))
        del keep, keepalive
        return out

    def rolling_least_squares(self, b: Batch, kw: L.RollingKwargs, mode: int):
        fr, keep, dtype, memspace, n, G = self._frame(b)
        F = len(b.features) + (1 if b.add_intercept else 0)
        shape = (n, F) if mode == L.COEFFICIENTS else (n,)
        o, v, m = self._alloc_out(memspace, shape, True, b.target.values)
        if mode == L.COEFFICIENTS:
            L.check(self._lib.b200ols_rolling_least_squares_coefficients(self._ctx, C.byref(fr), C.byref(kw), C.byref(o)))
        else:
            L.check(self._lib.b200ols_rolling_least_squares(self._ctx, C.byref(fr), C.byref(kw), mode, C.byref(o)))
        del keep
        return v, m


    # -- fused multi-GPU gather (peer memory over NVLink) -------------------------------------------------
    def device_alloc(self, nbytes: int) -> int:
        p = self._lib.b200ols_device_alloc(self._ctx, nbytes)
        if not p:
            raise L.B200OLSError(L.ERR_CUDA, self._lib.b200ols_last_error().decode())
        return p

    def zero_device(self, ptr: int, nbytes: int):
        L.check(self._lib.b200ols_device_memset(self._ctx, C.c_void_p(ptr), 0, nbytes))

    def device_free(self, ptr: int):
        self._lib.b200ols_device_free(self._ctx, C.c_void_p(ptr))

    def ipc_export(self, ptr: int) -> bytes:
        h = (C.c_uint8 * 64)()
        L.check(self._lib.b200ols_ipc_export(self._ctx, C.c_void_p(ptr), h))
        return bytes(h)

    def ipc_open(self, handle: bytes) -> int:
        h = (C.c_uint8 * 64).from_buffer_copy(handle)
        out = C.c_void_p()
        L.check(self._lib.b200ols_ipc_open(self._ctx, h, C.byref(out)))
        return out.value

    def ipc_close(self, ptr: int):
        L.check(self._lib.b200ols_ipc_close(self._ctx, C.c_void_p(ptr)))

    def copy_to_host(self, dst: np.ndarray, src_ptr: int):
        L.check(self._lib.b200ols_copy_to_host(self._ctx, dst.ctypes.data, C.c_void_p(src_ptr), dst.nbytes))

    def set_peer_gather(self, peer_ptrs: Sequence[int], group_base: int, total_groups: int):
        arr = (C.c_void_p * max(len(peer_ptrs), 1))(*peer_ptrs)
        L.check(self._lib.b200ols_set_peer_gather(self._ctx, len(peer_ptrs), arr, group_base, total_groups))

    def set_peer_flags(self, flag_ptrs: Sequence[int], rank: int):
        arr = (C.c_void_p * max(len(flag_ptrs), 1))(*flag_ptrs)
        L.check(self._lib.b200ols_set_peer_flags(self._ctx, len(flag_ptrs), arr, rank))

    def peer_step_complete(self, step: int):
        rc = self._lib.b200ols_peer_step_complete(self._ctx, step)
        if rc != 0:
            L.check(rc)

    def peer_step_signal(self, step: int):
        rc = self._lib.b200ols_peer_step_signal(self._ctx, step)
        if rc != 0:
            L.check(rc)

    def peer_step_signal_wait(self, signal_step: int, wait_step: int):
        rc = self._lib.b200ols_peer_step_signal_wait(self._ctx, signal_step, wait_step)
        if rc != 0:
            L.check(rc)

    def peer_arm_step(self, signal_step: int, wait_step: int):
        rc = self._lib.b200ols_peer_arm_step(self._ctx, signal_step, wait_step)
        if rc != 0:
            L.check(rc)

    def peer_step_wait(self, step: int):
        rc = self._lib.b200ols_peer_step_wait(self._ctx, step)
        if rc != 0:
            L.check(rc)

    def peer_timed_out(self) -> bool:
        return self._lib.b200ols_peer_timed_out(self._ctx) != 0

    def predict(self, coefficients: List[Col], features: List[Col], add_intercept: bool, null_policy: int):
        """b200ols_predict: row-wise dot of features with per-row coefficient columns."""
        n = len(coefficients[0])
        dev = coefficients[0].is_device
        for c_ in coefficients:
            if dev:
                if c_.values.dtype != torch.float64:
                    c_.values = c_.values.to(torch.float64)
            elif c_.values.dtype != np.float64 or not c_.values.flags.c_contiguous:
                c_.values = np.ascontiguousarray(c_.values, dtype=np.float64)
        f32 = len(features) > 0 and all((f.values.dtype == (torch.float32 if dev else np.float32)) for f in features)
        for f in features:
            if dev:
                want = torch.float32 if f32 else torch.float64
                if f.values.dtype != want:
                    f.values = f.values.to(want)
            else:
                want = np.float32 if f32 else np.float64
                if f.values.dtype != want or not f.values.flags.c_contiguous:
                    f.values = np.ascontiguousarray(f.values, dtype=want)
        cc = (L.Column * len(coefficients))(*[L.Column(_ptr(c_.values), _ptr(c_.validity)) for c_ in coefficients])
        ff = (L.Column * max(len(features), 1))(*[L.Column(_ptr(f.values), _ptr(f.validity)) for f in features])
        memspace = L.DEVICE if dev else L.HOST
        o, v, m = self._alloc_out(memspace, (n,), True, coefficients[0].values)
        L.check(self._lib.b200ols_predict(self._ctx, n, len(coefficients), L.F32 if f32 else L.F64, memspace, cc, ff,
                                          1 if add_intercept else 0, null_policy, C.byref(o)))
        return v, m


_engines = {}

CUDA_STREAM_LEGACY = 1  # cudaStreamLegacy: the handle of the (legacy) default stream


def get_engine(device: int = 0, torch_stream: bool = False) -> Engine:
    """process-wide default engines.  Host frames use the engine's own stream; frames of CUDA torch tensors
    run on torch's current stream so that they are ordered with the producer of those tensors."""
    stream = None
    if torch_stream and torch is not None:
        stream = torch.cuda.current_stream(device).cuda_stream or CUDA_STREAM_LEGACY
    key = (device, stream)
    e = _engines.get(key)
    if e is None:
        e = _engines[key] = Engine(device, stream)
    return e


def nan_or(x: Optional[float]) -> float:
    return math.nan if x is None else float(x)

"""Plays the polars side of the plugin FFI against libb200ols.so's `_polars_plugin_*` symbols (tests only).

polars is not installed here, so pyarrow provides the Arrow C Data Interface halves: input Series are exported
chunk by chunk into `SeriesExport` structs (polars-ffi version_0 layout), the returned
`SeriesExport` is imported back the way polars-ffi's `import_series` does (the ArrowArray is MOVED out, then the
export's release callback frees the containers), and the ownership protocol is checked on the way."""
import ctypes as C
import pickle

import pyarrow as pa

from polars_ols_b200 import _lib as L


class ArrowSchema(C.Structure):
    pass


class ArrowArray(C.Structure):
    pass


ArrowSchema._fields_ = [("format", C.c_char_p), ("name", C.c_char_p), ("metadata", C.c_char_p), ("flags", C.c_int64),
                        ("n_children", C.c_int64), ("children", C.POINTER(C.POINTER(ArrowSchema))),
                        ("dictionary", C.POINTER(ArrowSchema)), ("release", C.c_void_p), ("private_data", C.c_void_p)]
ArrowArray._fields_ = [("length", C.c_int64), ("null_count", C.c_int64), ("offset", C.c_int64), ("n_buffers", C.c_int64),
                       ("n_children", C.c_int64), ("buffers", C.POINTER(C.c_void_p)),
                       ("children", C.POINTER(C.POINTER(ArrowArray))), ("dictionary", C.POINTER(ArrowArray)),
                       ("release", C.c_void_p), ("private_data", C.c_void_p)]


class SeriesExport(C.Structure):
    pass


RELEASE_FN = C.CFUNCTYPE(None, C.POINTER(SeriesExport))
SeriesExport._fields_ = [("field", C.POINTER(ArrowSchema)), ("arrays", C.POINTER(C.POINTER(ArrowArray))), ("len", C.c_size_t),
                         ("release", C.c_void_p), ("private_data", C.c_void_p)]
SCHEMA_RELEASE = C.CFUNCTYPE(None, C.POINTER(ArrowSchema))


class Exported:
    """keeps the C structs of the exported inputs alive and records what the callee released"""

    def __init__(self, series):
        self.n = len(series)
        self.exports = (SeriesExport * self.n)()
        self.keep = []
        self.series_released = [0] * self.n
        self.arrays = []
        for i, (name, data) in enumerate(series):
            chunks = data.chunks if isinstance(data, pa.ChunkedArray) else [data]
            schema = ArrowSchema()
            pa.field(name, chunks[0].type)._export_to_c(C.addressof(schema))
            arrs = [ArrowArray() for _ in chunks]
            for a, ch in zip(arrs, chunks):
                ch._export_to_c(C.addressof(a))
            ptrs = (C.POINTER(ArrowArray) * len(arrs))(*[C.pointer(a) for a in arrs])

            def make_release(idx, schema=schema):
                def rel(e):
                    self.series_released[idx] += 1
                    if schema.release:                      # frees the schema, NOT the (moved-out) arrays
                        SCHEMA_RELEASE(schema.release)(C.pointer(schema))
                    e.contents.release = None
                return RELEASE_FN(rel)

            fn = make_release(i)
            self.exports[i].field = C.pointer(schema)
            self.exports[i].arrays = ptrs
            self.exports[i].len = len(arrs)
            self.exports[i].release = C.cast(fn, C.c_void_p)
            self.keep += [schema, arrs, ptrs, fn, chunks]
            self.arrays.append(arrs)

    def check_consumed(self):
        """every chunk array released exactly once (release pointer cleared by its producer), every series released once"""
        assert self.series_released == [1] * self.n, self.series_released
        for arrs in self.arrays:
            for a in arrs:
                assert not a.release, "an input ArrowArray was not released by the callee"


def import_series(ret: SeriesExport):
    """polars-ffi import_series: move the arrays out, import, then call the export's release."""
    assert ret.release, "return_value was not filled"
    name = ret.field.contents.name.decode()
    assert ret.len == 1
    moved = ArrowArray()
    C.memmove(C.addressof(moved), ret.arrays[0], C.sizeof(ArrowArray))      # std::ptr::read
    arr = pa.Array._import_from_c(C.addressof(moved), C.addressof(ret.field.contents))
    RELEASE_FN(ret.release)(C.pointer(ret))                                  # frees containers ("drop the box, not the array")
    assert not ret.release
    return name, arr


def call(fn_name: str, series, kwargs: dict):
    """series: [(name, pyarrow Array | ChunkedArray)], kwargs pickled like polars' register_plugin_function does.
    Returns (name, pyarrow array) or raises RuntimeError(last error message)."""
    lib = L.load()
    fn = getattr(lib, f"_polars_plugin_{fn_name}")
    fn.restype = None
    fn.argtypes = [C.POINTER(SeriesExport), C.c_size_t, C.c_char_p, C.c_size_t, C.POINTER(SeriesExport)]
    lib._polars_plugin_get_last_error_message.restype = C.c_char_p
    ex = Exported(series)
    blob = pickle.dumps(kwargs, protocol=5)
    ret = SeriesExport()
    fn(ex.exports, ex.n, blob, len(blob), C.byref(ret))
    ex.check_consumed()
    if not ret.release:
        raise RuntimeError(lib._polars_plugin_get_last_error_message().decode())
    return import_series(ret)


def call_field(fn_name: str, fields):
    """fields: [pa.Field] -> pa.Field returned by _polars_plugin_field_<fn>"""
    lib = L.load()
    fn = getattr(lib, f"_polars_plugin_field_{fn_name}")
    fn.restype = None
    fn.argtypes = [C.POINTER(ArrowSchema), C.c_size_t, C.POINTER(ArrowSchema)]
    arr = (ArrowSchema * len(fields))()
    for i, f in enumerate(fields):
        f._export_to_c(C.addressof(arr[i]))
    ret = ArrowSchema()
    fn(arr, len(fields), C.byref(ret))
    for i in range(len(fields)):                                             # the caller keeps ownership of its fields
        if arr[i].release:
            SCHEMA_RELEASE(arr[i].release)(C.pointer(arr[i]))
    if not ret.release:
        lib._polars_plugin_get_last_error_message.restype = C.c_char_p
        raise RuntimeError(lib._polars_plugin_get_last_error_message().decode())
    return pa.Field._import_from_c(C.addressof(ret))

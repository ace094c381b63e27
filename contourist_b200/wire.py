"""Wire-format text through the C ABI (`ctr_wire_format`, include/contourist_b200.h).

The reference builds its JSON / HTML payloads as "[" + sep.join(str(x) ...) + "]" (html_demo.py:118-161,
morph_geometry.py:91-128); `format_rows` produces the same bytes from a 2D numpy array on host threads.
Like every other entry point there is no Python fallback: without the library this raises.
"""
import ctypes

import numpy as np

from . import engine as E

_DTYPES = {np.dtype(np.int32): 0, np.dtype(np.int64): 1, np.dtype(np.uint32): 2, np.dtype(np.float32): 3,
           np.dtype(np.float64): 4}


def _prototype(lib):
    if getattr(lib, "_wire_bound", False):
        return
    c = ctypes
    lib.ctr_wire_format.argtypes = [c.c_void_p, c.c_int, c.c_int64, c.c_int64, c.c_char_p, c.c_char_p, c.c_char_p, c.c_char_p,
                                    c.c_int, c.POINTER(c.c_void_p), c.POINTER(c.c_int64)]
    lib.ctr_wire_format.restype = c.c_int
    lib.ctr_wire_free.argtypes = [c.c_void_p]
    lib.ctr_wire_free.restype = None
    lib._wire_bound = True


def format_rows(array, row_prefix="", col_sep=",", row_suffix="", row_sep=",\n", threads=0):
    """str: "[" + row_sep.join(row_prefix + col_sep.join(str(v) for v in row) + row_suffix for row in array) + "]"
    with str() of Python ints / floats (float32 widened first).  array: [rows, cols] or [rows] (cols = 1)."""
    a = np.asarray(array)
    if a.ndim == 1:
        a = a.reshape(-1, 1)
    if a.ndim != 2:
        raise ValueError("format_rows wants a 1D or 2D array, got shape %r" % (a.shape,))
    if a.dtype not in _DTYPES:
        if np.issubdtype(a.dtype, np.integer) or a.dtype == np.bool_:
            a = a.astype(np.int64)
        elif np.issubdtype(a.dtype, np.floating):
            a = a.astype(np.float64)
        else:
            raise ValueError("format_rows: unsupported dtype %r" % (a.dtype,))
    a = np.ascontiguousarray(a)
    lib = E.load_library()
    _prototype(lib)
    out, n = ctypes.c_void_p(), ctypes.c_int64()
    rc = lib.ctr_wire_format(a.ctypes.data_as(ctypes.c_void_p) if a.size else None, _DTYPES[a.dtype], a.shape[0], a.shape[1],
                             row_prefix.encode(), col_sep.encode(), row_suffix.encode(), row_sep.encode(), int(threads),
                             ctypes.byref(out), ctypes.byref(n))
    if rc != 0:
        raise E.EngineError("ctr_wire_format failed (%d)" % rc)
    try:
        return ctypes.string_at(out.value, n.value).decode("ascii")
    finally:
        lib.ctr_wire_free(out)

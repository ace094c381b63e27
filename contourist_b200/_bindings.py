"""ctypes prototypes for the 2D / 4D / post-processing entry points."""
import ctypes


class Mt2dParams(ctypes.Structure):
    _fields_ = [("field", ctypes.c_void_p), ("dtype", ctypes.c_int32), ("flags", ctypes.c_uint32),
                ("n0", ctypes.c_int64), ("n1", ctypes.c_int64), ("levels", ctypes.POINTER(ctypes.c_double)),
                ("nlevels", ctypes.c_int32), ("reserved", ctypes.c_int32),
                ("origin", ctypes.c_double * 2), ("delta", ctypes.c_double * 2),
                ("i_lo", ctypes.c_int64), ("i_hi", ctypes.c_int64), ("row_offset", ctypes.c_int64)]


class Mt2dCounts(ctypes.Structure):
    _fields_ = [("n_segments", ctypes.c_int64), ("n_active_squares", ctypes.c_int64),
                ("fmin", ctypes.c_double), ("fmax", ctypes.c_double)]


def bind(lib):
    vp, i32 = ctypes.c_void_p, ctypes.c_int
    lib.ctr_mt2d_run.argtypes = [vp, ctypes.POINTER(Mt2dParams), ctypes.POINTER(Mt2dCounts)]
    lib.ctr_mt2d_run.restype = i32
    lib.ctr_mt2d_fetch.argtypes = [vp, vp, vp, vp]
    lib.ctr_mt2d_fetch.restype = i32

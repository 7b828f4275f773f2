"""TEST INFRASTRUCTURE: ctypes front-end of the CPU oracle (oracle/fgoicp_oracle.c).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import
this module.  The product package (fast_go_icp_b200) never does.
"""
import ctypes as C
import os

import numpy as np

from . import build_oracle

_lib = None

_f32p = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
_u64p = np.ctypeslib.ndpointer(dtype=np.uint64, flags="C_CONTIGUOUS")


def lib():
    global _lib
    if _lib is not None:
        return _lib
    path = build_oracle.OUT
    if not os.path.exists(path) or os.path.getmtime(path) < os.path.getmtime(build_oracle.SRC):
        build_oracle.build()
    L = C.CDLL(path)
    L.orc_set_modes.argtypes = [C.c_int, C.c_int]
    L.orc_set_sin_table.argtypes = [_f32p, _f32p, C.c_int]
    L.orc_num_threads.restype = C.c_int
    L.orc_set_num_threads.argtypes = [C.c_int]
    L.orc_set_nn_mode.argtypes = [C.c_int]
    L.orc_set_lut_mode.argtypes = [C.c_int]
    L.orc_rotation.argtypes = [C.c_float, C.c_float, C.c_float, _f32p]
    L.orc_rotation.restype = C.c_float
    L.orc_overlaps_so3.argtypes = [C.c_float] * 4
    L.orc_overlaps_so3.restype = C.c_int
    L.orc_in_so3.argtypes = [C.c_float] * 3
    L.orc_in_so3.restype = C.c_int
    L.orc_rot_sin.argtypes = [C.c_float]
    L.orc_rot_sin.restype = C.c_float
    L.orc_center.argtypes = [_f32p, C.c_size_t, _f32p]
    L.orc_scaling_factor.argtypes = [_f32p, C.c_size_t]
    L.orc_scaling_factor.restype = C.c_float
    L.orc_scale.argtypes = [_f32p, C.c_size_t, C.c_float]
    L.orc_ranges.argtypes = [_f32p, C.c_size_t, _f32p, _f32p]
    L.orc_restore_translation.argtypes = [_f32p, _f32p, C.c_float, _f32p, _f32p, _f32p]
    L.orc_lut_dims.argtypes = [_f32p, _f32p, C.c_float, _i32p]
    L.orc_lut_build.argtypes = [_f32p, C.c_size_t, _f32p, C.c_float, _i32p, _f32p]
    L.orc_lut_sample.argtypes = [_f32p, _i32p, _f32p, C.c_float, _f32p, C.c_size_t, _f32p]
    L.orc_bounds.argtypes = [_f32p, _i32p, _f32p, C.c_float, _f32p, C.c_size_t, _f32p, C.c_float,
                             C.c_int, _f32p, C.c_int, _f32p, _f32p]
    L.orc_nn.argtypes = [_f32p, C.c_size_t, _f32p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_int,
                         C.c_void_p, C.c_void_p]
    L.orc_sse.argtypes = [_f32p, C.c_size_t, _f32p, C.c_size_t, _f32p, _f32p]
    L.orc_sse.restype = C.c_float
    L.orc_closest_orthogonal.argtypes = [_f32p, _f32p]
    L.orc_icp.argtypes = [_f32p, C.c_size_t, _f32p, C.c_size_t, C.c_int, C.c_float, _f32p, _f32p,
                          _f32p, _f32p, C.POINTER(C.c_int)]
    L.orc_icp.restype = C.c_float
    L.orc_bnb_r3.argtypes = [_f32p, C.c_size_t, _f32p, C.c_size_t, _f32p, _i32p, _f32p, C.c_float,
                             _f32p, C.c_int, C.c_float, C.c_float, C.c_int, _f32p,
                             C.POINTER(C.c_uint64), C.POINTER(C.c_uint32)]
    L.orc_bnb_r3.restype = C.c_float
    L.orc_run.argtypes = [_f32p, C.c_size_t, _f32p, C.c_size_t, _f32p, _i32p, _f32p, C.c_float,
                          C.c_float, _f32p, _f32p, _u64p]
    L.orc_run.restype = C.c_float
    L.orc_set_trim_k.argtypes = [C.c_size_t]
    L.orc_get_trim_k.restype = C.c_size_t
    L.orc_trim_count.argtypes = [C.c_size_t, C.c_float]
    L.orc_trim_count.restype = C.c_size_t
    _lib = L
    return L


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def set_modes(weight_mode=0, interp_mode=0):
    lib().orc_set_modes(int(weight_mode), int(interp_mode))


def trim_count(ns, rho):
    """Inliers kept for a trim fraction: ns - floor(float32(ns) * rho)."""
    return int(lib().orc_trim_count(int(ns), float(rho)))


def set_trim_k(k):
    """Trimmed registration (extension, off by default): every sum over the data points keeps the k smallest
    residuals only.  0 switches it off.  Global state: reset it when done (tests use the `trimmed` context manager)."""
    lib().orc_set_trim_k(int(k))


class trimmed:
    def __init__(self, k):
        self.k = int(k)

    def __enter__(self):
        set_trim_k(self.k)
        return self

    def __exit__(self, *a):
        set_trim_k(0)


def set_sin_table(spans, vals):
    s, v = _f32(spans), _f32(vals)
    lib().orc_set_sin_table(s, v, len(s))


def num_threads():
    return lib().orc_num_threads()


def set_num_threads(n):
    lib().orc_set_num_threads(int(n))


def rotation(x, y, z):
    R = np.zeros(9, np.float32)
    r = lib().orc_rotation(x, y, z, R)
    return R, r


def overlaps_so3(x, y, z, span):
    return bool(lib().orc_overlaps_so3(x, y, z, span))


def in_so3(x, y, z):
    return bool(lib().orc_in_so3(x, y, z))


def rot_sin(span):
    return lib().orc_rot_sin(span)


def preprocess(model, data):
    """FastGoICP ctor preprocessing (fgoicp.hpp:13-19): centre source, centre target, scale both by
    the SOURCE's 1/max|coord|, bbox of the target.  Returns a dict."""
    L = lib()
    pcs = _f32(data).copy()
    pct = _f32(model).copy()
    off_s = np.zeros(3, np.float32)
    off_t = np.zeros(3, np.float32)
    L.orc_center(pcs, len(pcs), off_s)
    L.orc_center(pct, len(pct), off_t)
    s = L.orc_scaling_factor(pcs, len(pcs))
    L.orc_scale(pcs, len(pcs), s)
    L.orc_scale(pct, len(pct), s)
    mn = np.zeros(3, np.float32)
    mx = np.zeros(3, np.float32)
    L.orc_ranges(pct, len(pct), mn, mx)
    return dict(model=pct, data=pcs, offset_pcs=off_s, offset_pct=off_t, scale=s, bbox_min=mn, bbox_max=mx)


def restore_translation(R, t, s, offset_pcs, offset_pct):
    out = np.zeros(3, np.float32)
    lib().orc_restore_translation(_f32(R), _f32(t), s, _f32(offset_pcs), _f32(offset_pct), out)
    return out


def lut_dims(bbox_min, bbox_max, res):
    d = np.zeros(3, np.int32)
    lib().orc_lut_dims(_f32(bbox_min), _f32(bbox_max), res, d)
    return d


def lut_build(model, bbox_min, bbox_max, res):
    model = _f32(model)
    dims = lut_dims(bbox_min, bbox_max, res)
    out = np.zeros(int(dims[0]) * int(dims[1]) * int(dims[2]), np.float32)
    lib().orc_lut_build(model, len(model), _f32(bbox_min), res, dims, out)
    return out, dims


def lut_sample(lut, dims, bbox_min, res, q):
    q = _f32(q)
    out = np.zeros(len(q), np.float32)
    lib().orc_lut_sample(_f32(lut), np.ascontiguousarray(dims, np.int32), _f32(bbox_min), res, q, len(q), out)
    return out


def bounds(lut, dims, bbox_min, res, data, R, rot_span, fix_rot, tcubes):
    data = _f32(data)
    tc = _f32(tcubes).reshape(-1, 4)
    T = len(tc)
    lb = np.zeros(T, np.float32)
    ub = np.zeros(T, np.float32)
    lib().orc_bounds(_f32(lut), np.ascontiguousarray(dims, np.int32), _f32(bbox_min), res, data,
                     len(data), _f32(R), rot_span, int(bool(fix_rot)), tc, T, lb, ub)
    return lb, ub


def set_nn_mode(mode):
    """0: brute-force scans (the reference's kernels, literally); 1 (default): exact k-d tree, same winners."""
    lib().orc_set_nn_mode(int(mode))


def set_lut_mode(mode):
    """0: every grid node scans every model point (the reference's kernel, literally); 1 (default): k-d tree, same values."""
    lib().orc_set_lut_mode(int(mode))


def nn(model, q, R=None, t=None, rooted=False):
    model, q = _f32(model), _f32(q)
    idx = np.zeros(len(q), np.int32)
    d2 = np.zeros(len(q), np.float32)
    Rp = _f32(R).ctypes.data if R is not None else None
    tp = _f32(t).ctypes.data if t is not None else None
    # keep temporaries alive across the call
    Rk = _f32(R) if R is not None else None
    tk = _f32(t) if t is not None else None
    lib().orc_nn(model, len(model), q, len(q), Rk.ctypes.data if Rk is not None else None,
                 tk.ctypes.data if tk is not None else None, int(bool(rooted)),
                 idx.ctypes.data, d2.ctypes.data)
    del Rp, tp
    return idx, d2


def sse(model, data, R, t):
    model, data = _f32(model), _f32(data)
    return lib().orc_sse(model, len(model), data, len(data), _f32(R), _f32(t))


def closest_orthogonal(ABt):
    out = np.zeros(9, np.float32)
    lib().orc_closest_orthogonal(_f32(ABt).reshape(9), out)
    return out


def icp(model, data, max_iter, thr, R0, t0):
    model, data = _f32(model), _f32(data)
    R = np.zeros(9, np.float32)
    t = np.zeros(3, np.float32)
    it = C.c_int(0)
    e = lib().orc_icp(model, len(model), data, len(data), int(max_iter), thr, _f32(R0).reshape(9),
                      _f32(t0), R, t, C.byref(it))
    return e, R, t, it.value


def bnb_r3(model, data, lut, dims, bbox_min, res, rot_xyz_span, fix_rot, best_sse, sse_threshold, batch=32):
    model, data = _f32(model), _f32(data)
    bt = np.zeros(3, np.float32)
    ev = C.c_uint64(0)
    nb = C.c_uint32(0)
    ub = lib().orc_bnb_r3(model, len(model), data, len(data), _f32(lut),
                          np.ascontiguousarray(dims, np.int32), _f32(bbox_min), res,
                          _f32(rot_xyz_span), int(bool(fix_rot)), best_sse, sse_threshold, int(batch),
                          bt, C.byref(ev), C.byref(nb))
    return ub, bt, ev.value, nb.value


def run(model, data, lut, dims, bbox_min, res, mse_threshold):
    """Best-first Go-ICP on centred+scaled clouds; returns (sse, R, t, stats)."""
    model, data = _f32(model), _f32(data)
    R = np.zeros(9, np.float32)
    t = np.zeros(3, np.float32)
    stats = np.zeros(4, np.uint64)
    e = lib().orc_run(model, len(model), data, len(data), _f32(lut), np.ascontiguousarray(dims, np.int32),
                      _f32(bbox_min), res, mse_threshold, R, t, stats)
    return e, R, t, dict(cubes=int(stats[0]), icps=int(stats[1]), evals=int(stats[2]), batches=int(stats[3]))

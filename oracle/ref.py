"""TEST INFRASTRUCTURE: ctypes front-end of oracle/_ref/libfgoicp_ref.so -- the UNMODIFIED reference
sources compiled by oracle/build_ref.py.  Needs a GPU (the reference is CUDA-only).  Only tests and
bench.py's reference arm import this."""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "_ref", "libfgoicp_ref.so")
_f32p = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
_lib = None


def available():
    return os.path.exists(LIB_PATH)


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(LIB_PATH)
        vp = C.c_void_p
        L.ref_create.argtypes = [_f32p, C.c_size_t, _f32p, C.c_size_t, C.c_float, C.c_float]
        L.ref_create.restype = vp
        L.ref_destroy.argtypes = [vp]
        L.ref_get_preprocessed.argtypes = [vp, _f32p, _f32p, _f32p, _f32p, C.POINTER(C.c_float), _f32p, _f32p]
        L.ref_lut_dims.argtypes = [vp, _i32p]
        L.ref_lut_download.argtypes = [vp, _f32p]
        L.ref_lut_sample.argtypes = [vp, _f32p, C.c_int, _f32p]
        L.ref_bounds.argtypes = [vp, _f32p, C.c_int, _f32p, C.c_int, _f32p, _f32p]
        L.ref_sse.argtypes = [vp, _f32p, _f32p]
        L.ref_sse.restype = C.c_float
        L.ref_icp.argtypes = [vp, _f32p, _f32p, C.c_int, C.c_float, _f32p, _f32p]
        L.ref_icp.restype = C.c_float
        L.ref_bnb_r3.argtypes = [vp, _f32p, C.c_int, C.c_float, _f32p]
        L.ref_bnb_r3.restype = C.c_float
        L.ref_run.argtypes = [vp, _f32p, _f32p, _f32p, _f32p]
        L.ref_run.restype = C.c_float
        L.ref_run_trace.argtypes = [vp, _f32p, _f32p, _f32p, C.c_int, C.POINTER(C.c_int)]
        L.ref_run_trace.restype = C.c_float
        L.ref_rot_sin.argtypes = [_f32p, C.c_int, _f32p]
        L.ref_sse_threshold.argtypes = [vp]
        L.ref_sse_threshold.restype = C.c_float
        _lib = L
    return _lib


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def rot_sin(spans):
    """sin(span * sqrt3 * pi / 2) from a kernel compiled inside the reference build with the reference's constants and
    flags (registration.cu:41-42): the values its bound kernel uses, obtained without the library under test."""
    s = _f32(spans).ravel()
    out = np.zeros(len(s), np.float32)
    assert lib().ref_rot_sin(s, len(s), out) == 0
    return out


class Reference:
    """icp::FastGoICP of the reference (constructor does centring, scaling and the brute-force LUT build)."""

    def __init__(self, target, source, lut_resolution, mse_threshold):
        t, s = _f32(target).reshape(-1, 3), _f32(source).reshape(-1, 3)
        self.nt, self.ns = len(t), len(s)
        self._h = lib().ref_create(t, len(t), s, len(s), lut_resolution, mse_threshold)

    def close(self):
        if self._h:
            lib().ref_destroy(self._h)
            self._h = None

    def preprocessed(self):
        model = np.zeros((self.nt, 3), np.float32)
        data = np.zeros((self.ns, 3), np.float32)
        off_s, off_t = np.zeros(3, np.float32), np.zeros(3, np.float32)
        mn, mx = np.zeros(3, np.float32), np.zeros(3, np.float32)
        s = C.c_float(0)
        lib().ref_get_preprocessed(self._h, model, data, off_s, off_t, C.byref(s), mn, mx)
        return dict(model=model, data=data, offset_pcs=off_s, offset_pct=off_t, scale=np.float32(s.value),
                    bbox_min=mn, bbox_max=mx)

    def lut(self):
        dims = np.zeros(3, np.int32)
        lib().ref_lut_dims(self._h, dims)
        out = np.zeros(int(dims[0]) * int(dims[1]) * int(dims[2]), np.float32)
        rc = lib().ref_lut_download(self._h, out)
        assert rc == 0
        return out, dims

    def lut_sample(self, q):
        q = _f32(q).reshape(-1, 3)
        out = np.zeros(len(q), np.float32)
        assert lib().ref_lut_sample(self._h, q, len(q), out) == 0
        return out

    def bounds(self, rot_xyz_span, fix_rot, tcubes):
        tc = _f32(tcubes).reshape(-1, 4)
        lb, ub = np.zeros(len(tc), np.float32), np.zeros(len(tc), np.float32)
        lib().ref_bounds(self._h, _f32(rot_xyz_span), int(bool(fix_rot)), tc, len(tc), lb, ub)
        return lb, ub

    def sse(self, R, t):
        return lib().ref_sse(self._h, _f32(R).reshape(9), _f32(t))

    def icp(self, R0, t0, max_iter, thr):
        R, t = np.zeros(9, np.float32), np.zeros(3, np.float32)
        e = lib().ref_icp(self._h, _f32(R0).reshape(9), _f32(t0), int(max_iter), thr, R, t)
        return e, R, t

    def bnb_r3(self, rot_xyz_span, fix_rot, best_sse):
        bt = np.zeros(3, np.float32)
        ub = lib().ref_bnb_r3(self._h, _f32(rot_xyz_span), int(bool(fix_rot)), best_sse, bt)
        return ub, bt

    def run(self):
        R, t, Rn, tn = (np.zeros(9, np.float32), np.zeros(3, np.float32), np.zeros(9, np.float32),
                        np.zeros(3, np.float32))
        sse = lib().ref_run(self._h, R, t, Rn, tn)
        return sse, R, t, Rn, tn

    def run_trace(self, cap=1 << 16):
        """run() with the Debug log on: returns (sse, R, t, best error printed after every refinement)."""
        R, t, tr = np.zeros(9, np.float32), np.zeros(3, np.float32), np.zeros(cap, np.float32)
        n = C.c_int(0)
        sse = lib().ref_run_trace(self._h, R, t, tr, cap, C.byref(n))
        return sse, R, t, tr[:min(n.value, cap)].copy()

    def sse_threshold(self):
        return lib().ref_sse_threshold(self._h)

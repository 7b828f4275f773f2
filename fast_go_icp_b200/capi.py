"""ctypes binding of the C ABI in include/fgoicp_c.h (libfgoicp_b200.so).

Thin by design: plain numpy arrays in, numpy arrays out, every non-zero status raised as
FgoicpError with the library's message.  There is no CPU fallback: if the CUDA library is missing
or no GPU is present, calls fail loudly.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
# FGOICP_LIB selects an experimental build of the same library (scripts/build_variant.py)
LIB_PATH = os.environ.get("FGOICP_LIB") or os.path.join(HERE, "libfgoicp_b200.so")

SAMPLER_GRID, SAMPLER_PACKED, SAMPLER_TEX = 0, 1, 2
BUILD_PACKED, BUILD_TEX, BUILD_BRUTE_LUT, BUILD_KEEP_ORDER = 1, 2, 4, 8
BUILD_DEFAULT = BUILD_PACKED | BUILD_TEX

EXPORTS = [
    "fgoicp_last_error", "fgoicp_version", "fgoicp_ctx_create", "fgoicp_ctx_destroy", "fgoicp_ctx_info",
    "fgoicp_set_sampler", "fgoicp_set_trim", "fgoicp_set_stream", "fgoicp_set_nn_mode", "fgoicp_set_phased", "fgoicp_set_bnb_mode", "fgoicp_set_icp_mode", "fgoicp_gather_probe", "fgoicp_lut_download", "fgoicp_lut_sample", "fgoicp_rot_sin",
    "fgoicp_bounds_batch", "fgoicp_bounds_multi", "fgoicp_bounds_multi_dev", "fgoicp_sse", "fgoicp_nn",
    "fgoicp_icp", "fgoicp_bnb_r3", "fgoicp_bnb_r3_batch", "fgoicp_so3_level_ub", "fgoicp_so3_level_lb",
    "fgoicp_preprocess", "fgoicp_preprocess_dev", "fgoicp_icp_batch",
]
PRE_REFERENCE, PRE_TREE_CENTROID, PRE_SCALE_BOTH = 0, 1, 2


class FgoicpError(RuntimeError):
    pass


class Info(C.Structure):
    _fields_ = [("nt", C.c_uint64), ("ns", C.c_uint64), ("dims", C.c_int32 * 3), ("resolution", C.c_float),
                ("scale", C.c_float), ("offset", C.c_float * 3), ("grid_bytes", C.c_uint64),
                ("packed_bytes", C.c_uint64), ("device", C.c_int32), ("sm_count", C.c_int32),
                ("sampler", C.c_int32), ("has_packed", C.c_int32), ("has_tex", C.c_int32),
                ("build_ms", C.c_float)]


class LevelStats(C.Structure):
    _fields_ = [("evals", C.c_uint64), ("n_icp", C.c_uint32), ("icp_iters", C.c_uint32),
                ("ms_bnb_ub", C.c_float), ("ms_icp", C.c_float), ("ms_bnb_lb", C.c_float),
                ("best_icp_index", C.c_int32)]


class Normalisation(C.Structure):
    _fields_ = [("offset_pcs", C.c_float * 3), ("offset_pct", C.c_float * 3), ("scale", C.c_float),
                ("bbox_min", C.c_float * 3), ("bbox_max", C.c_float * 3), ("device_ms", C.c_float)]


_f32p = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
_lib = None


def lib():
    """Loads libfgoicp_b200.so (built by fast_go_icp_b200/build.py); raises if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise FgoicpError("CUDA library %s is missing: run `python fast_go_icp_b200/build.py` "
                          "(there is no CPU fallback)" % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    L.fgoicp_last_error.restype = C.c_char_p
    L.fgoicp_version.restype = C.c_char_p
    vp = C.c_void_p
    L.fgoicp_ctx_create.argtypes = [_f32p, C.c_size_t, _f32p, C.c_size_t, _f32p, _f32p, C.c_float, C.c_int,
                                    C.c_uint, C.POINTER(vp)]
    L.fgoicp_ctx_destroy.argtypes = [vp]
    L.fgoicp_ctx_info.argtypes = [vp, C.POINTER(Info)]
    L.fgoicp_set_sampler.argtypes = [vp, C.c_int]
    L.fgoicp_set_stream.argtypes = [vp, vp]
    L.fgoicp_set_nn_mode.argtypes = [vp, C.c_int]
    L.fgoicp_set_phased.argtypes = [vp, C.c_int]
    L.fgoicp_set_bnb_mode.argtypes = [vp, C.c_int]
    L.fgoicp_set_icp_mode.argtypes = [vp, C.c_int]
    L.fgoicp_set_trim.argtypes = [vp, C.c_float, C.POINTER(C.c_uint64)]
    L.fgoicp_gather_probe.argtypes = [vp, C.c_size_t, C.c_int, C.c_int, C.POINTER(C.c_float)]
    L.fgoicp_lut_download.argtypes = [vp, _f32p, C.c_size_t]
    L.fgoicp_lut_sample.argtypes = [vp, _f32p, C.c_size_t, C.c_int, _f32p]
    L.fgoicp_rot_sin.argtypes = [vp, _f32p, C.c_int, _f32p]
    L.fgoicp_bounds_batch.argtypes = [vp, _f32p, C.c_float, C.c_int, _f32p, C.c_int, _f32p, _f32p]
    L.fgoicp_bounds_multi.argtypes = [vp, _f32p, C.c_int, C.c_int, _f32p, C.c_int, _f32p, _f32p]
    L.fgoicp_bounds_multi_dev.argtypes = [vp, vp, C.c_int, C.c_int, vp, C.c_int, vp, vp, vp]
    L.fgoicp_sse.argtypes = [vp, _f32p, _f32p, C.POINTER(C.c_float)]
    L.fgoicp_nn.argtypes = [vp, _f32p, _f32p, C.c_int, vp, vp]
    L.fgoicp_icp.argtypes = [vp, _f32p, _f32p, C.c_int, C.c_float, C.POINTER(C.c_float), _f32p, _f32p,
                             C.POINTER(C.c_int)]
    L.fgoicp_icp_batch.argtypes = [vp, _f32p, _f32p, C.c_int, C.c_int, C.c_float, _f32p, _f32p, _f32p, vp]
    L.fgoicp_bnb_r3.argtypes = [vp, _f32p, C.c_int, C.c_float, C.c_float, C.POINTER(C.c_float), _f32p,
                                C.POINTER(C.c_uint64)]
    L.fgoicp_bnb_r3_batch.argtypes = [vp, _f32p, C.c_int, C.c_int, C.c_float, C.c_float, _f32p, _f32p, vp]
    L.fgoicp_so3_level_ub.argtypes = [vp, vp, C.c_int, C.c_float, C.c_float, vp, vp,
                                      C.POINTER(C.c_float), _f32p, _f32p, C.POINTER(LevelStats)]
    L.fgoicp_so3_level_lb.argtypes = [vp, vp, C.c_int, C.c_float, C.c_float, vp, C.POINTER(LevelStats)]
    L.fgoicp_preprocess.argtypes = [_f32p, C.c_size_t, _f32p, C.c_size_t, C.c_int, C.c_uint, C.POINTER(Normalisation)]
    L.fgoicp_preprocess_dev.argtypes = [vp, C.c_size_t, vp, C.c_size_t, C.c_int, C.c_uint, vp, C.POINTER(Normalisation)]
    for name in EXPORTS:
        getattr(L, name).restype = getattr(L, name).restype if name in ("fgoicp_last_error", "fgoicp_version") else C.c_int
    _lib = L
    return L


def _check(rc, what):
    if rc != 0:
        raise FgoicpError("%s failed (%d): %s" % (what, rc, lib().fgoicp_last_error().decode()))


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _normalisation_dict(model, data, n):
    F = np.float32
    return dict(model=model, data=data, offset_pcs=np.array(n.offset_pcs, F), offset_pct=np.array(n.offset_pct, F),
                scale=F(n.scale), bbox_min=np.array(n.bbox_min, F), bbox_max=np.array(n.bbox_max, F),
                device_ms=float(n.device_ms))


def preprocess(target, source, device=0, flags=PRE_REFERENCE):
    """FastGoICP constructor preprocessing on the GPU (fgoicp/fgoicp.hpp:13-19, fgoicp.cpp:176-287) through
    fgoicp_preprocess: centre both clouds, scale by 1 / max|coord| of the source, range of the target.
    Same keys as driver.preprocess; flags = PRE_REFERENCE is bit-identical to it."""
    model = _f32(target).reshape(-1, 3).copy()
    data = _f32(source).reshape(-1, 3).copy()
    n = Normalisation()
    _check(lib().fgoicp_preprocess(model, len(model), data, len(data), int(device), int(flags), C.byref(n)),
           "fgoicp_preprocess")
    return _normalisation_dict(model, data, n)


def preprocess_dev(d_model_ptr, nt, d_data_ptr, ns, device=0, flags=PRE_REFERENCE, cuda_stream_ptr=None):
    """Same on clouds already in HBM (raw device pointers to n*3 floats, modified in place)."""
    n = Normalisation()
    _check(lib().fgoicp_preprocess_dev(C.c_void_p(d_model_ptr), int(nt), C.c_void_p(d_data_ptr), int(ns), int(device),
                                       int(flags), C.c_void_p(cuda_stream_ptr or 0), C.byref(n)), "fgoicp_preprocess_dev")
    return _normalisation_dict(None, None, n)


class Context:
    """Owns one fgoicp_ctx: both clouds (already centred and scaled) and the NN grid on one GPU."""

    def __init__(self, model, data, bbox_min, bbox_max, lut_resolution, device=0, flags=BUILD_DEFAULT):
        L = lib()
        self._h = C.c_void_p()
        model, data = _f32(model).reshape(-1, 3), _f32(data).reshape(-1, 3)
        self.ns, self.nt = len(data), len(model)
        _check(L.fgoicp_ctx_create(model, len(model), data, len(data), _f32(bbox_min), _f32(bbox_max),
                                   float(lut_resolution), int(device), int(flags), C.byref(self._h)),
               "fgoicp_ctx_create")

    def close(self):
        if self._h:
            lib().fgoicp_ctx_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def info(self):
        i = Info()
        _check(lib().fgoicp_ctx_info(self._h, C.byref(i)), "fgoicp_ctx_info")
        return i

    def set_sampler(self, sampler):
        _check(lib().fgoicp_set_sampler(self._h, int(sampler)), "fgoicp_set_sampler")

    def set_trim(self, trim_fraction):
        """Trimmed registration (extension; 0 = off = the reference).  Returns the number of inliers kept."""
        k = C.c_uint64(0)
        _check(lib().fgoicp_set_trim(self._h, float(trim_fraction), C.byref(k)), "fgoicp_set_trim")
        return int(k.value)

    def set_bnb_mode(self, mode):
        """0: automatic, 1: persistent per-cube searches, 2: round-synchronous searches (same results)."""
        _check(lib().fgoicp_set_bnb_mode(self._h, int(mode)), "fgoicp_set_bnb_mode")

    def set_icp_mode(self, mode):
        """0: automatic (default), 1: one launch per stage and iteration, 2: persistent loop kernel (same results)."""
        _check(lib().fgoicp_set_icp_mode(self._h, int(mode)), "fgoicp_set_icp_mode")

    def set_phased(self, on):
        _check(lib().fgoicp_set_phased(self._h, int(bool(on))), "fgoicp_set_phased")

    def set_nn_mode(self, mode):
        _check(lib().fgoicp_set_nn_mode(self._h, int(mode)), "fgoicp_set_nn_mode")

    def gather_probe(self, nbytes, width_bytes, blocks_per_sm=4):
        out = C.c_float(0)
        _check(lib().fgoicp_gather_probe(self._h, int(nbytes), int(width_bytes), int(blocks_per_sm), C.byref(out)),
               "fgoicp_gather_probe")
        return out.value

    def set_stream(self, cuda_stream_ptr):
        _check(lib().fgoicp_set_stream(self._h, C.c_void_p(cuda_stream_ptr)), "fgoicp_set_stream")

    def lut_download(self):
        i = self.info()
        dims = np.array(list(i.dims), np.int32)
        out = np.zeros(int(dims[0]) * int(dims[1]) * int(dims[2]), np.float32)
        _check(lib().fgoicp_lut_download(self._h, out, out.size), "fgoicp_lut_download")
        return out, dims

    def lut_sample(self, q, sampler):
        q = _f32(q).reshape(-1, 3)
        out = np.zeros(len(q), np.float32)
        _check(lib().fgoicp_lut_sample(self._h, q, len(q), int(sampler), out), "fgoicp_lut_sample")
        return out

    def rot_sin(self, spans):
        s = _f32(spans)
        out = np.zeros(len(s), np.float32)
        _check(lib().fgoicp_rot_sin(self._h, s, len(s), out), "fgoicp_rot_sin")
        return out

    def bounds_batch(self, R, rot_span, fix_rot, tcubes):
        tc = _f32(tcubes).reshape(-1, 4)
        lb = np.zeros(len(tc), np.float32)
        ub = np.zeros(len(tc), np.float32)
        _check(lib().fgoicp_bounds_batch(self._h, _f32(R).reshape(9), float(rot_span), int(bool(fix_rot)), tc,
                                         len(tc), lb, ub), "fgoicp_bounds_batch")
        return lb, ub

    def bounds_multi(self, rot_cubes, fix_rot, tcubes):
        rc = _f32(rot_cubes).reshape(-1, 4)
        tc = _f32(tcubes).reshape(len(rc), -1, 4)
        T = tc.shape[1]
        lb = np.zeros((len(rc), T), np.float32)
        ub = np.zeros((len(rc), T), np.float32)
        _check(lib().fgoicp_bounds_multi(self._h, rc, len(rc), int(bool(fix_rot)), tc, T, lb, ub),
               "fgoicp_bounds_multi")
        return lb, ub

    def bounds_multi_dev(self, d_rot, Rn, fix_rot, d_tc, T, d_lb, d_ub, d_best_ub=None):
        """Device-pointer form: arguments are integer device addresses (e.g. torch tensor.data_ptr())."""
        _check(lib().fgoicp_bounds_multi_dev(self._h, C.c_void_p(d_rot), int(Rn), int(bool(fix_rot)),
                                             C.c_void_p(d_tc), int(T), C.c_void_p(d_lb), C.c_void_p(d_ub),
                                             C.c_void_p(d_best_ub) if d_best_ub else None),
               "fgoicp_bounds_multi_dev")

    def sse(self, R, t):
        out = C.c_float(0)
        _check(lib().fgoicp_sse(self._h, _f32(R).reshape(9), _f32(t), C.byref(out)), "fgoicp_sse")
        return out.value

    def nn(self, R, t, rooted=False):
        idx = np.zeros(self.ns, np.int32)
        d2 = np.zeros(self.ns, np.float32)
        _check(lib().fgoicp_nn(self._h, _f32(R).reshape(9), _f32(t), int(bool(rooted)), idx.ctypes.data,
                               d2.ctypes.data), "fgoicp_nn")
        return idx, d2

    def icp(self, R0, t0, max_iter, thr):
        R = np.zeros(9, np.float32)
        t = np.zeros(3, np.float32)
        e = C.c_float(0)
        it = C.c_int(0)
        _check(lib().fgoicp_icp(self._h, _f32(R0).reshape(9), _f32(t0), int(max_iter), float(thr), C.byref(e),
                                R, t, C.byref(it)), "fgoicp_icp")
        return e.value, R, t, it.value

    def icp_batch(self, R0s, t0s, max_iter, thr):
        """n refinements at once (fgoicp_icp_batch): returns sse[n], R[n][9], t[n][3], iters[n]."""
        R0s, t0s = _f32(R0s).reshape(-1, 9), _f32(t0s).reshape(-1, 3)
        n = len(R0s)
        e, R, t = np.zeros(n, np.float32), np.zeros((n, 9), np.float32), np.zeros((n, 3), np.float32)
        it = np.zeros(n, np.int32)
        _check(lib().fgoicp_icp_batch(self._h, R0s, t0s, n, int(max_iter), float(thr), e, R, t, it.ctypes.data),
               "fgoicp_icp_batch")
        return e, R, t, it

    def bnb_r3(self, rot_cube, fix_rot, best_sse, sse_threshold):
        bt = np.zeros(3, np.float32)
        ub = C.c_float(0)
        ev = C.c_uint64(0)
        _check(lib().fgoicp_bnb_r3(self._h, _f32(rot_cube), int(bool(fix_rot)), float(best_sse),
                                   float(sse_threshold), C.byref(ub), bt, C.byref(ev)), "fgoicp_bnb_r3")
        return ub.value, bt, ev.value

    def bnb_r3_batch(self, rot_cubes, fix_rot, best_sse, sse_threshold):
        rc = _f32(rot_cubes).reshape(-1, 4)
        ub = np.zeros(len(rc), np.float32)
        bt = np.zeros((len(rc), 3), np.float32)
        ev = np.zeros(len(rc), np.uint64)
        _check(lib().fgoicp_bnb_r3_batch(self._h, rc, len(rc), int(bool(fix_rot)), float(best_sse),
                                         float(sse_threshold), ub, bt, ev.ctypes.data), "fgoicp_bnb_r3_batch")
        return ub, bt, ev

    def so3_level_ub(self, cubes, best_sse, sse_threshold, best_R, best_t):
        cubes = _f32(cubes).reshape(-1, 4)
        n = len(cubes)
        ub = np.zeros(n, np.float32)
        bt = np.zeros((n, 3), np.float32)
        io_sse = C.c_float(best_sse)
        R = _f32(best_R).reshape(9).copy()
        t = _f32(best_t).copy()
        st = LevelStats()
        _check(lib().fgoicp_so3_level_ub(self._h, cubes.ctypes.data if n else None, n, float(best_sse),
                                         float(sse_threshold), ub.ctypes.data if n else None,
                                         bt.ctypes.data if n else None, C.byref(io_sse), R, t, C.byref(st)),
               "fgoicp_so3_level_ub")
        return ub, bt, io_sse.value, R, t, st

    def so3_level_lb(self, cubes, best_sse, sse_threshold):
        cubes = _f32(cubes).reshape(-1, 4)
        n = len(cubes)
        lb = np.zeros(n, np.float32)
        st = LevelStats()
        _check(lib().fgoicp_so3_level_lb(self._h, cubes.ctypes.data if n else None, n, float(best_sse),
                                         float(sse_threshold), lb.ctypes.data if n else None, C.byref(st)),
               "fgoicp_so3_level_lb")
        return lb, st

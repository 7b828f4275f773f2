"""Test stand-in for fast_go_icp_b200.capi.Context backed by the CPU oracle: same methods the driver
uses, so the level-synchronous / sharded host logic can be exercised without a GPU (tests only)."""
import types

import numpy as np

from oracle import oracle as O


class OracleContext:
    def __init__(self, model, data, bbox_min, bbox_max, lut_resolution, device=0, flags=0):
        self.model, self.data = np.ascontiguousarray(model, np.float32), np.ascontiguousarray(data, np.float32)
        self.bbox_min, self.res = np.asarray(bbox_min, np.float32), float(lut_resolution)
        self.lut, self.dims = O.lut_build(self.model, bbox_min, bbox_max, self.res)
        self.ns, self.nt = len(self.data), len(self.model)

    def info(self):
        return types.SimpleNamespace(build_ms=0.0, dims=list(self.dims))

    def set_sampler(self, s):
        pass

    def set_trim(self, trim_fraction):
        """Trimmed registration: the oracle's global switch stays on for the life of this context (tests only)."""
        k = O.trim_count(self.ns, trim_fraction)
        O.set_trim_k(0 if k == self.ns else k)
        return k

    def close(self):
        O.set_trim_k(0)

    def icp(self, R0, t0, max_iter, thr):
        return O.icp(self.model, self.data, max_iter, thr, R0, t0)

    def _bnb(self, cube, fix_rot, best_sse, thr):
        return O.bnb_r3(self.model, self.data, self.lut, self.dims, self.bbox_min, self.res, cube, fix_rot, best_sse, thr)

    def so3_level_ub(self, cubes, best_sse, thr, best_R, best_t):
        cubes = np.asarray(cubes, np.float32).reshape(-1, 4)
        n = len(cubes)
        ub, bt = np.zeros(n, np.float32), np.zeros((n, 3), np.float32)
        st = types.SimpleNamespace(evals=0, n_icp=0, icp_iters=0, ms_bnb_ub=0.0, ms_icp=0.0, ms_bnb_lb=0.0, best_icp_index=-1)
        io, R, t = np.float32(best_sse), np.array(best_R, np.float32), np.array(best_t, np.float32)
        for i, c in enumerate(cubes):
            ub[i], bt[i], ev, _ = self._bnb(c, True, best_sse, thr)
            st.evals += ev
        for i, c in enumerate(cubes):
            if not (float(ub[i]) < float(best_sse) * 1.8):
                continue
            R0, _ = O.rotation(*c[:3])
            e, Ri, ti, it = O.icp(self.model, self.data, 100, 0.005, R0, bt[i])
            st.n_icp += 1
            st.icp_iters += it
            if e < io:
                io, R, t, st.best_icp_index = np.float32(e), Ri, ti, i
        return ub, bt, float(io), R, t, st

    def so3_level_lb(self, cubes, best_sse, thr):
        cubes = np.asarray(cubes, np.float32).reshape(-1, 4)
        lb = np.zeros(len(cubes), np.float32)
        st = types.SimpleNamespace(evals=0, ms_bnb_lb=0.0)
        for i, c in enumerate(cubes):
            lb[i], _, ev, _ = self._bnb(c, False, best_sse, thr)
            st.evals += ev
        return lb, st

"""Host-side mirror of icp::FastGoICP over the C ABI, with the per-level rotation frontier sharded
across ranks (one process per GPU, torch.distributed for the exchange).

Same algorithm as fast_go_icp_b200/csrc/fgoicp_host.cpp (level-synchronous form of the reference's
fgoicp/fgoicp.cpp:10-100); the C++ class is the drop-in for src/main.cpp, this module is what
bench.py and the multi-GPU tests drive.  All host arithmetic that the reference does in fp32 on the
host (preprocessing, Rotation, overlaps_SO3, restore_translation) is done here in np.float32 with
the same operation order, so C++ and Python drivers agree bit for bit.

Sharding: at each level the children to evaluate are dealt round-robin to the
ranks; after the fixed-rotation phase the best SSE is MIN-reduced (all_reduce on a packed
(sse bits, child index) key -- NCCL over NVLink on GPUs) and the winner's pose broadcast; per-cube
results are all-gathered so that every rank rebuilds the identical next frontier.  The SCHEDULE is independent
of the number of ranks by construction (which cube is searched against which incumbent, which cube is refined);
the values are too, except that a bound is an fp64 sum of fp32 terms rounded once and the order of its partial
sums follows the launch geometry (cluster size depends on the number of cubes a rank holds): a sum that sits within
~1e-13 relative of an fp32 rounding boundary can round differently.  Never observed (same SSE bits at
1, 2, 3, 8 ranks); the tests compare bits and would say so.
"""
import time

import numpy as np

from . import capi

F = np.float32
M_INF = F(1e10)


def rotation_matrix(x, y, z):
    """Rotation(x, y, z) of the reference (fgoicp/common.hpp:37-57) in fp32.
    Returns (R as 9 floats column-major, r)."""
    x, y, z = F(x), F(y), F(z)
    r = F(F(F(x * x) + F(y * y)) + F(z * z))
    R = np.array([1, 0, 0, 0, 1, 0, 0, 0, 1], F)
    if r > F(1.0):
        return R, r
    ww = F(F(1.0) - r)
    w = F(np.sqrt(ww))
    wx, xx = F(w * x), F(x * x)
    wy, xy, yy = F(w * y), F(x * y), F(y * y)
    wz, xz, yz, zz = F(w * z), F(x * z), F(y * z), F(z * z)
    two = F(2.0)
    R = np.array([
        F(F(F(ww + xx) - yy) - zz), F(two * F(xy - wz)), F(two * F(xz + wy)),
        F(two * F(xy + wz)), F(F(F(ww - xx) + yy) - zz), F(two * F(yz - wx)),
        F(two * F(xz - wy)), F(two * F(yz + wx)), F(F(F(ww - xx) - yy) + zz)], F)
    return R, F(np.sqrt(r))


def in_so3(x, y, z):
    return rotation_matrix(x, y, z)[1] <= F(1.0)


def overlaps_so3(x, y, z, span):
    """RotNode::overlaps_SO3 (fgoicp/common.hpp:99-103)."""
    x, y, z, span = F(x), F(y), F(z), F(span)
    r = rotation_matrix(x, y, z)[1]
    a = F(F(abs(x) + abs(y)) + abs(z))
    v = F(F(r - F(F(F(2.0) * span) * a)) + F(F(F(3.0) * span) * span))
    return v <= F(1.0)


def preprocess(target, source):
    """FastGoICP constructor preprocessing (fgoicp/fgoicp.hpp:13-19, fgoicp.cpp:176-287)."""
    pcs = np.ascontiguousarray(source, F).reshape(-1, 3).copy()
    pct = np.ascontiguousarray(target, F).reshape(-1, 3).copy()

    def centre(pc):
        c = np.cumsum(pc, axis=0, dtype=F)[-1] / F(len(pc))     # sequential fp32 sum, index order
        pc -= c
        return -c

    off_s = centre(pcs)
    off_t = centre(pct)
    s = F(1.0) / np.abs(pcs).max()
    pcs *= s
    pct *= s
    return dict(model=pct, data=pcs, offset_pcs=off_s.astype(F), offset_pct=off_t.astype(F), scale=F(s),
                bbox_min=pct.min(axis=0), bbox_max=pct.max(axis=0))


def restore_translation(R, t, s, offset_pcs, offset_pct):
    """fgoicp/fgoicp.hpp:87-90: t / s + R * offset_pcs - offset_pct (fp32, left to right)."""
    R = np.asarray(R, F).reshape(9)
    out = np.zeros(3, F)
    for r in range(3):
        rp = F(F(F(R[r] * offset_pcs[0]) + F(R[3 + r] * offset_pcs[1])) + F(R[6 + r] * offset_pcs[2]))
        out[r] = F(F(F(t[r]) / s + rp) - offset_pct[r])
    return out


class _Comm:
    """Minimal collective layer: torch.distributed when a process group is up, identity otherwise."""

    def __init__(self, group=None):
        self.dist = None
        self.rank, self.world = 0, 1
        try:
            import torch.distributed as dist
            if dist.is_available() and dist.is_initialized():
                self.dist = dist
                self.group = group
                self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        except ImportError:
            pass
        if self.dist is not None:
            import torch
            self.torch = torch
            self.dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" \
                else torch.device("cpu")

    def min_key(self, key):
        """MIN all-reduce of one int64."""
        if self.dist is None:
            return key
        t = self.torch.tensor([key], dtype=self.torch.int64, device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MIN, group=self.group)
        return int(t.item())

    def bcast_f32(self, arr, src):
        if self.dist is None:
            return arr
        t = self.torch.from_numpy(np.ascontiguousarray(arr, F)).to(self.dev)
        self.dist.broadcast(t, src=src, group=self.group)
        return t.cpu().numpy()

    def _buffers(self, n):
        """Staging for the exchange, allocated once and grown geometrically: pinned host buffers on both sides of the
        device buffers NCCL works on (gloo: plain host tensors), so that a wave's exchange is two async copies, one
        collective and ONE stream synchronisation -- no allocation, no implicit synchronising copies."""
        if getattr(self, "_cap", 0) < n:
            cap = max(4096, 2 * n)
            t = self.torch
            cuda = self.dev.type == "cuda"
            self._h_in = t.empty(cap, dtype=t.float32, pin_memory=cuda)
            self._h_out = t.empty(cap * self.world, dtype=t.float32, pin_memory=cuda)
            self._d_in = t.empty(cap, dtype=t.float32, device=self.dev) if cuda else self._h_in
            self._d_out = t.empty(cap * self.world, dtype=t.float32, device=self.dev) if cuda else self._h_out
            self._cap = cap
        return self._h_in, self._d_in, self._d_out, self._h_out

    def exchange(self, header, local, n_total):
        """ONE all-gather per wave: every rank contributes a fixed-size header (uint32 words) and its round-robin
        shard of rows (float32).  Returns (headers [world, H] uint32, full rows [n_total, k]) on every rank."""
        header = np.ascontiguousarray(header, np.uint32)
        if self.dist is None:
            return header[None, :], local
        k = local.shape[1]
        per = (n_total + self.world - 1) // self.world
        n = len(header) + per * k
        h_in, d_in, d_out, h_out = self._buffers(n)
        buf = h_in.numpy()[:n]
        buf[:len(header)] = header.view(F)                      # bit patterns travel unchanged (copies only)
        buf[len(header):len(header) + local.size] = local.ravel()
        buf[len(header) + local.size:] = 0
        if self.dev.type == "cuda":
            d_in[:n].copy_(h_in[:n], non_blocking=True)
            self.dist.all_gather_into_tensor(d_out[:self.world * n], d_in[:n], group=self.group)
            h_out[:self.world * n].copy_(d_out[:self.world * n], non_blocking=True)
            self.torch.cuda.current_stream().synchronize()
        else:
            self.dist.all_gather_into_tensor(d_out[:self.world * n], d_in[:n], group=self.group)
        allb = h_out.numpy()[:self.world * n].reshape(self.world, n)
        heads = np.ascontiguousarray(allb[:, :len(header)]).view(np.uint32)
        full = np.empty((per * self.world, k), F)
        for r in range(self.world):
            full[r::self.world] = allb[r, len(header):].reshape(per, k)
        return heads, full[:n_total]

    def gather_rows(self, local, n_total):
        """local: rows of this rank's round-robin shard (global rows rank, rank+world, ...).
        Returns the full (n_total, k) array on every rank (one all-gather, one device-to-host copy)."""
        if self.dist is None:
            return local
        return self.exchange(np.zeros(1, np.uint32), local, n_total)[1]


class FastGoICP:
    """Python mirror of icp::FastGoICP (reference fgoicp/fgoicp.hpp:10-108)."""

    def __init__(self, target, source, lut_resolution, mse_threshold, device=0, sampler=None,
                 flags=capi.BUILD_PACKED, group=None, ctx_factory=None, wave1=32, skip_dead_lb=True,
                 schedule="level", trim_fraction=0.0, device_preprocess=False, preprocess_flags=0):
        t0 = time.perf_counter()
        # device_preprocess: centring / scaling / ranges on the GPU through fgoicp_preprocess;
        # with preprocess_flags = 0 bit-identical to the host pass below
        if device_preprocess:
            self.pp = capi.preprocess(target, source, device=device, flags=preprocess_flags)
        else:
            self.pp = preprocess(target, source)
        self.ns, self.nt = len(self.pp["data"]), len(self.pp["model"])
        self.mse_threshold = F(mse_threshold)
        self.sse_threshold = F(F(self.ns) * self.mse_threshold)            # fgoicp.hpp:23
        # ctx_factory: tests inject a stand-in for the CUDA context to exercise the sharding logic on CPU
        make = ctx_factory if ctx_factory is not None else capi.Context
        self.ctx = make(self.pp["model"], self.pp["data"], self.pp["bbox_min"], self.pp["bbox_max"],
                        lut_resolution, device=device, flags=flags)
        if sampler is not None:
            self.ctx.set_sampler(sampler)
        # trimmed registration (extension, off by default): sums run over the n_inliers smallest residuals, and the
        # convergence threshold scales with them (Go-ICP: SSEThresh = MSEThresh * inlierNum)
        self.n_inliers = self.ns
        if trim_fraction and trim_fraction > 0:
            self.n_inliers = self.ctx.set_trim(trim_fraction)
            self.sse_threshold = F(F(self.n_inliers) * self.mse_threshold)
        self.comm = _Comm(group)
        # "level": level-synchronous, whole levels in flight, shardable over ranks (default);
        # "bestfirst": the reference's own visiting order (fgoicp.cpp:32-100), one rotation cube at a time
        assert schedule in ("level", "bestfirst")
        self.schedule = schedule
        self.wave1 = int(wave1)      # size of the first wave of a level (0: whole level at once)
        self.skip_dead_lb = bool(skip_dead_lb)
        self.trace = None            # best-first schedule: set to [] to record (cube, ub, icp sse, best sse) per refinement
        self.best_sse = M_INF
        self.best_R = np.array([1, 0, 0, 0, 1, 0, 0, 0, 1], F)
        self.best_t = np.zeros(3, F)
        self.stats = dict(bound_evals=0, rot_cubes=0, icp_runs=0, icp_iters=0, levels=0,
                          ctor_ms=(time.perf_counter() - t0) * 1e3, lut_build_ms=self.ctx.info().build_ms,
                          ms_bnb_ub=0.0, ms_icp=0.0, ms_bnb_lb=0.0, level_log=[],
                          ms_calls=0.0, ms_exchange=0.0, exchanges=0)

    def close(self):
        self.ctx.close()

    def get_best_error(self):
        return float(self.best_sse)

    def get_best_transform(self):
        return self.best_R.copy(), self.best_t.copy()

    def run(self):
        """Returns (R 3x3 with y = R @ x + t, t) mapping source -> target in original coordinates."""
        t0 = time.perf_counter()
        I = np.array([1, 0, 0, 0, 1, 0, 0, 0, 1], F)
        e, _, _, it = self.ctx.icp(I, np.zeros(3, F), 100, 0.05)           # fgoicp.cpp:12-14
        self.stats["ms_first_icp"] = (time.perf_counter() - t0) * 1e3
        self.best_sse = F(e)
        self.stats["icp_runs"] += 1
        self.stats["icp_iters"] += it
        self.stats["initial_icp_sse"] = float(e)
        if self.schedule == "bestfirst":
            self._search_best_first()
        else:
            self._search_level_synchronous()
        t1 = time.perf_counter()
        self.stats["ms_search_wall"] = (t1 - t0) * 1e3 - self.stats["ms_first_icp"]
        e, R, t, it = self.ctx.icp(self.best_R, self.best_t, 100, 0.0005)  # fgoicp.cpp:22-23
        self.stats["ms_final_icp"] = (time.perf_counter() - t1) * 1e3
        self.best_sse, self.best_R, self.best_t = F(e), R, t
        self.stats["icp_runs"] += 1
        self.stats["icp_iters"] += it
        self.stats["run_ms"] = (time.perf_counter() - t0) * 1e3
        t_out = restore_translation(R, t, self.pp["scale"], self.pp["offset_pcs"], self.pp["offset_pct"])
        return R.reshape(3, 3).T.copy(), t_out

    # -- outer search, reference order -------------------------------------------------------------
    def _search_best_first(self):
        """branch_and_bound_SO3 exactly as the reference walks it (fgoicp.cpp:32-100): a best-first heap of
        rotation cubes (smaller lb first, then larger span; equal keys in insertion order -- the total order the
        oracle uses for std::priority_queue's unspecified ties), each popped cube split in 8, every child searched
        on its own.  Single rank only; this is the schedule the parity tests compare with the reference's run()."""
        import heapq
        assert self.comm.world == 1, "the best-first schedule does not shard"
        thr = self.sse_threshold
        heap, seq = [(F(0.0), -F(1.0), 0, (F(0), F(0), F(0), F(1.0), F(0.0), F(self.best_sse)))], 1
        s = self.stats
        while heap:
            nlb, _, _, (x, y, z, pspan, _, nub) = heapq.heappop(heap)
            if F(self.best_sse) - nlb <= thr:                               # fgoicp.cpp:44
                break
            span = F(pspan / F(2.0))
            for k in range(8):
                if span < F(0.05):                                          # fgoicp.cpp:53
                    continue
                cx = F(F(x - span) + F(F(k & 1) * pspan))
                cy = F(F(y - span) + F(F((k >> 1) & 1) * pspan))
                cz = F(F(z - span) + F(F((k >> 2) & 1) * pspan))
                if not overlaps_so3(cx, cy, cz, span):                      # fgoicp.cpp:61
                    continue
                if not in_so3(cx, cy, cz):                                  # fgoicp.cpp:62-66
                    heapq.heappush(heap, (nlb, -span, seq, (cx, cy, cz, span, nlb, nub)))
                    seq += 1
                    continue
                cube = np.array([cx, cy, cz, span], F)
                ub, bt, ev = self.ctx.bnb_r3(cube, True, self.best_sse, thr)              # fgoicp.cpp:69
                s["bound_evals"] += int(ev)
                s["rot_cubes"] += 1
                if float(ub) < float(self.best_sse) * 1.8:                  # fgoicp.cpp:74 (double compare)
                    R0, _ = rotation_matrix(cx, cy, cz)
                    e, R, t, it = self.ctx.icp(R0, bt, 100, 0.005)          # fgoicp.cpp:76
                    s["icp_runs"] += 1
                    s["icp_iters"] += it
                    if e < self.best_sse:                                   # fgoicp.cpp:79-84
                        self.best_sse, self.best_R, self.best_t = F(e), R, t
                    if self.trace is not None:                              # what the reference logs at fgoicp.cpp:85
                        self.trace.append((float(cx), float(cy), float(cz), float(span), float(ub), float(e), float(self.best_sse)))
                lb, _, ev = self.ctx.bnb_r3(cube, False, self.best_sse, thr)              # fgoicp.cpp:90
                s["bound_evals"] += int(ev)
                if lb >= self.best_sse:                                     # fgoicp.cpp:92
                    continue
                heapq.heappush(heap, (F(lb), -span, seq, (cx, cy, cz, span, F(lb), F(ub))))
                seq += 1

    # -- outer search, level-synchronous -------------------------------------------------------------
    def _search_level_synchronous(self):
        comm = self.comm
        thr = self.sse_threshold
        # frontier rows: x, y, z, span, lb, ub
        frontier = np.array([[0, 0, 0, 1, 0, self.best_sse]], F)
        while len(frontier):
            keep = ~(F(self.best_sse) - frontier[:, 4] <= thr)             # fgoicp.cpp:44
            frontier = frontier[keep]
            if not len(frontier):
                break
            pspan = frontier[0, 3]
            span = F(pspan / F(2.0))
            if span < F(0.05):                                             # fgoicp.cpp:53
                break
            # children of every open node, (node, octant) order, in fp32 array arithmetic (same operations and
            # order as the scalar code of the reference: x - span + bit * parent_span, fgoicp.cpp:55-59)
            bits = np.array([[(j >> a) & 1 for a in range(3)] for j in range(8)], F)          # (8, 3)
            ctr = (frontier[:, None, :3] - span) + bits[None, :, :] * pspan                     # (n, 8, 3) fp32
            cx, cy, cz = ctr[..., 0].ravel(), ctr[..., 1].ravel(), ctr[..., 2].ravel()
            rr = (cx * cx + cy * cy) + cz * cz
            rq = np.where(rr > F(1.0), rr, np.sqrt(rr)).astype(F)                               # Rotation::r (Q3)
            a = (np.abs(cx) + np.abs(cy)) + np.abs(cz)
            overl = ((rq - (F(2.0) * span) * a) + (F(3.0) * span) * span) <= F(1.0)             # fgoicp.cpp:61
            inside = rq <= F(1.0)                                                               # fgoicp.cpp:62
            rows = np.empty((len(cx), 6), F)
            rows[:, 0], rows[:, 1], rows[:, 2], rows[:, 3] = cx, cy, cz, span
            rows[:, 4] = np.repeat(frontier[:, 4], 8)
            rows[:, 5] = np.repeat(frontier[:, 5], 8)
            nxt = rows[overl & ~inside]
            ev = rows[overl & inside]
            n_ev = len(ev)
            # Fixed-rotation phase in (up to) two waves.  The children of the parents with the smallest
            # fixed-rotation error go first: whatever their ICPs find tightens best_sse -- and with it every
            # inner search -- for the (much larger) rest of the level.  The split depends only on the level's
            # global data, so results are still independent of the number of ranks.
            ub = np.zeros(n_ev, F)
            bt = np.zeros((n_ev, 3), F)
            by_parent_ub = np.argsort(ev[:, 5], kind="stable")
            waves, lo, size = [], 0, self.wave1
            while lo < n_ev:                                   # wave sizes wave1, 4*wave1, 16*wave1, ... then the rest
                hi = n_ev if (size <= 0 or len(waves) >= 3) else min(n_ev, lo + size)
                waves.append(np.sort(by_parent_ub[lo:hi]))
                lo, size = hi, size * 4
            tot = dict(evals=0, n_icp=0, icp_iters=0, ms_ub=0.0, ms_icp=0.0, local=0)
            for widx in waves:
                if not len(widx):
                    continue
                my_idx = widx[comm.rank::comm.world]
                tc0 = time.perf_counter()
                ub_l, bt_l, e_l, R_l, t_l, st = self.ctx.so3_level_ub(ev[my_idx, :4], self.best_sse, thr,
                                                                     self.best_R, self.best_t)
                tc1 = time.perf_counter()
                # ONE exchange per wave: header = (sse bits, global child index, pose of this rank's best ICP), rows =
                # (ub, best_t) of this rank's cubes.  Global best = MIN over (sse bits, child index): ties -> lowest
                # child index, which is what a single rank's ascending scan picks.
                head = np.zeros(14, np.uint32)
                if st.best_icp_index >= 0 and e_l < self.best_sse:
                    head[0] = np.array([e_l], F).view(np.uint32)[0]
                    head[1] = np.uint32(int(my_idx[st.best_icp_index]))
                    head[2:14] = np.concatenate([R_l, t_l]).astype(F).view(np.uint32)
                else:
                    head[0] = np.array([self.best_sse], F).view(np.uint32)[0]
                    head[1] = np.uint32(0xffffffff)
                heads, ubt_w = comm.exchange(head, np.concatenate([ub_l[:, None], bt_l], axis=1), len(widx))
                self.stats["ms_calls"] += (tc1 - tc0) * 1e3
                self.stats["ms_exchange"] += (time.perf_counter() - tc1) * 1e3
                self.stats["exchanges"] += 1
                keys = (heads[:, 0].astype(np.uint64) << np.uint64(32)) | heads[:, 1].astype(np.uint64)
                win = int(np.argmin(keys))
                if heads[win, 1] != np.uint32(0xffffffff):
                    pose = heads[win, 2:14].view(F)
                    self.best_sse = heads[win, 0:1].view(F)[0]
                    self.best_R, self.best_t = pose[:9].copy(), pose[9:].copy()
                ub[widx], bt[widx] = ubt_w[:, 0], ubt_w[:, 1:]
                tot["evals"] += int(st.evals); tot["n_icp"] += int(st.n_icp); tot["icp_iters"] += int(st.icp_iters)
                tot["ms_ub"] += st.ms_bnb_ub; tot["ms_icp"] += st.ms_icp; tot["local"] += len(my_idx)

            mine = ev[comm.rank::comm.world, :4]
            if self.skip_dead_lb and F(span / F(2.0)) < F(0.05):
                # Leaf level: the children of these cubes are never evaluated (fgoicp.cpp:53), so their lower
                # bounds can only feed the loop's exit test -- the reference computes them anyway (fgoicp.cpp:90);
                # skipping them changes no output (and every rank knows the zeros: nothing to exchange).
                lb, st2 = np.zeros(n_ev, F), capi.LevelStats()
            else:
                tc0 = time.perf_counter()
                lb_l, st2 = self.ctx.so3_level_lb(mine, self.best_sse, thr)
                tc1 = time.perf_counter()
                lb = comm.gather_rows(lb_l[:, None], n_ev)[:, 0] if n_ev else np.zeros(0, F)
                self.stats["ms_calls"] += (tc1 - tc0) * 1e3
                self.stats["ms_exchange"] += (time.perf_counter() - tc1) * 1e3
                self.stats["exchanges"] += 1

            surv = lb < self.best_sse                                      # fgoicp.cpp:92
            kept = ev[surv].copy()
            kept[:, 4] = lb[surv]
            kept[:, 5] = ub[surv]
            frontier = np.concatenate([nxt, kept], axis=0)

            s = self.stats
            s["levels"] += 1
            s["rot_cubes"] += len(mine)
            s["bound_evals"] += tot["evals"] + int(st2.evals)
            s["icp_runs"] += tot["n_icp"]
            s["icp_iters"] += tot["icp_iters"]
            s["ms_bnb_ub"] += tot["ms_ub"]
            s["ms_icp"] += tot["ms_icp"]
            s["ms_bnb_lb"] += st2.ms_bnb_lb
            s["level_log"].append(dict(span=float(span), cubes=n_ev, local_cubes=len(mine), icps=tot["n_icp"],
                                       evals=tot["evals"] + int(st2.evals), best_sse=float(self.best_sse),
                                       survivors=len(frontier), ms_ub=tot["ms_ub"], ms_icp=tot["ms_icp"],
                                       ms_lb=st2.ms_bnb_lb))

"""N2: trimmed registration (extension -- the reference parses `trim` and ignores it, src/utilities.hpp:94).

With trim fraction rho every sum over the data points keeps the K = ns - floor(ns * rho) smallest residuals:
per-cube bounds, exact SSE, the ICP's Procrustes step.  rho = 0 must reproduce the untrimmed results bit for bit.
CPU part: the oracle's trimmed sums against a numpy restatement (sort, take K, fp64 sum).  GPU part: the CUDA path
against the oracle (bounds 1 ulp, SSE bit-exact, ICP iteration counts, inner-search evaluation counts) and an
end-to-end run with 25 % gross outliers."""
import numpy as np
import pytest

from fast_go_icp_b200 import workloads
from oracle import oracle as O

ULP = 2.4e-7
F = np.float32


def _lut(pp):
    return pp["lut"], pp["dims"], pp["bbox_min"], float(pp["res"])


def test_trim_count_rule():
    assert O.trim_count(300, 0.0) == 300 and O.trim_count(300, 0.1) == 270 and O.trim_count(10000, 0.25) == 7500
    assert O.trim_count(7, 0.5) == 4 and O.trim_count(3, 0.99) == 1          # ns - floor(float32(ns) * rho), >= 1


def test_oracle_trimmed_sums_match_numpy(small_problem):
    pp = small_problem
    ns = len(pp["data"])
    K = O.trim_count(ns, 0.2)
    R, _ = O.rotation(0.2, -0.1, 0.1)
    tc = workloads.translation_cube_list(4, level=2, seed=3)
    for fix_rot in (True, False):
        per_u = np.zeros((ns, len(tc)), F)
        per_l = np.zeros((ns, len(tc)), F)
        for i in range(ns):                                                 # per-point terms = bounds of a 1-point cloud
            per_l[i], per_u[i] = O.bounds(*_lut(pp), pp["data"][i:i + 1], R, 0.125, fix_rot, tc)
        with O.trimmed(K):
            lb, ub = O.bounds(*_lut(pp), pp["data"], R, 0.125, fix_rot, tc)
        want_u = np.sort(per_u, axis=0)[:K].astype(np.float64).sum(axis=0).astype(F)
        want_l = np.sort(per_l, axis=0)[:K].astype(np.float64).sum(axis=0).astype(F)
        assert np.array_equal(ub, want_u) and np.array_equal(lb, want_l)
        lb0, ub0 = O.bounds(*_lut(pp), pp["data"], R, 0.125, fix_rot, tc)
        assert np.all(ub <= ub0) and np.all(lb <= lb0) and np.all(lb <= ub)
    t = F([0.02, 0.01, -0.03])
    _, d2 = O.nn(pp["model"], pp["data"], R, t, False)
    with O.trimmed(K):
        e = O.sse(pp["model"], pp["data"], R, t)
    assert e == F(np.sort(d2)[:K].astype(np.float64).sum())
    with O.trimmed(ns):                                                      # K = ns: trimming off, bit for bit
        assert O.sse(pp["model"], pp["data"], R, t) == O.sse(pp["model"], pp["data"], R, t)
    assert O.lib().orc_get_trim_k() == 0


def test_oracle_trimmed_icp_ignores_outliers(small_problem):
    pp = small_problem
    rng = np.random.default_rng(4)
    # a data cloud that IS registered at the identity (model points + small noise), then 20 % gross outliers
    data = (pp["model"][::6][:300] + rng.normal(scale=0.003, size=(300, 3))).astype(F)
    bad = rng.choice(len(data), 60, replace=False)
    data[bad] = rng.uniform(-1, 1, (60, 3)).astype(F)
    I = np.eye(3, dtype=F).ravel()
    Rs, _ = O.rotation(0.02, -0.015, 0.01)
    e_plain, R_plain, t_plain, _ = O.icp(pp["model"], data, 60, 0.001, Rs, np.zeros(3, F))
    with O.trimmed(O.trim_count(len(data), 0.25)):
        e_trim, R_trim, t_trim, _ = O.icp(pp["model"], data, 60, 0.001, Rs, np.zeros(3, F))
    assert e_trim < 0.05 * e_plain                                           # the outliers dominate the untrimmed SSE
    assert np.abs(R_trim - I).max() < 0.01 and np.abs(t_trim).max() < 0.01   # trimmed fit: back at the identity
    assert np.abs(R_trim - I).max() <= np.abs(R_plain - I).max() + 1e-6


# ---- GPU ------------------------------------------------------------------------------------------------------

@pytest.mark.gpu
def test_cuda_trimmed_operators_vs_oracle(small_problem, gpu_ctx):
    from fast_go_icp_b200 import capi
    pp = small_problem
    ns = len(pp["data"])
    rng = np.random.default_rng(6)
    for rho in (0.1, 0.37):
        K = gpu_ctx.set_trim(rho)
        assert K == O.trim_count(ns, rho)
        try:
            with O.trimmed(K):
                for sampler in (capi.SAMPLER_PACKED, capi.SAMPLER_GRID):
                    gpu_ctx.set_sampler(sampler)
                    for fix_rot in (True, False):
                        R, _ = O.rotation(0.2, -0.1, 0.1)
                        tc = workloads.translation_cube_list(37, level=3, seed=8)
                        lb, ub = gpu_ctx.bounds_batch(R, 0.125, fix_rot, tc)
                        wl, wu = O.bounds(*_lut(pp), pp["data"], R, 0.125, fix_rot, tc)
                        assert np.allclose(ub, wu, rtol=ULP, atol=0) and np.allclose(lb, wl, rtol=ULP, atol=0)
                gpu_ctx.set_sampler(capi.SAMPLER_PACKED)
                rot = workloads.rotation_cube_list(9, seed=2)
                tcs = np.stack([workloads.translation_cube_list(8, level=2, seed=20 + r) for r in range(9)])
                tcs[rng.random((9, 8)) < 0.3, 3] = -1.0                       # unused slots are skipped
                lb, ub = gpu_ctx.bounds_multi(rot, False, tcs)
                for r in range(9):
                    Rr, _ = O.rotation(*rot[r, :3])
                    live = tcs[r, :, 3] >= 0
                    wl, wu = O.bounds(*_lut(pp), pp["data"], Rr, float(rot[r, 3]), False, tcs[r][live])
                    assert np.allclose(ub[r][live], wu, rtol=ULP, atol=0) and np.allclose(lb[r][live], wl, rtol=ULP, atol=0)
                R, _ = O.rotation(0.05, 0.02, -0.04)
                t = F([0.02, 0.01, -0.03])
                assert gpu_ctx.sse(R, t) == O.sse(pp["model"], pp["data"], R, t)            # trimmed SSE, bit-exact
                e, Rg, tg, it = gpu_ctx.icp(R, t, 100, 0.005)
                we, wR, wt, wit = O.icp(pp["model"], pp["data"], 100, 0.005, R, t)
                assert it == wit and abs(e - we) <= 1e-6 * we
                assert np.allclose(Rg, wR, atol=2e-6) and np.allclose(tg, wt, atol=2e-6)
                thr = K * 1e-4
                cubes = np.float32([[0.25, -0.25, 0.25, 0.25], [0.0625, 0.1875, -0.0625, 0.0625]])
                for fix_rot in (True, False):
                    ubs, bts, evs = gpu_ctx.bnb_r3_batch(cubes, fix_rot, 1e10, thr)
                    for i, c in enumerate(cubes):
                        wub, wbt, wev, _ = O.bnb_r3(pp["model"], pp["data"], *_lut(pp), c, fix_rot, 1e10, thr)
                        assert evs[i] == wev and np.isclose(ubs[i], wub, rtol=ULP, atol=0) and np.array_equal(bts[i], wbt)
        finally:
            gpu_ctx.set_trim(0.0)
    # rho = 0 restores the untrimmed results bit for bit
    R, _ = O.rotation(0.05, 0.02, -0.04)
    assert gpu_ctx.sse(R, np.zeros(3, F)) == O.sse(pp["model"], pp["data"], R, np.zeros(3, F))


@pytest.mark.gpu
def test_cuda_trimmed_run_recovers_pose_despite_outliers():
    from fast_go_icp_b200 import driver
    w = workloads.synthetic_pair(nt=4000, ns=600, sigma=0.005, seed=13)
    rng = np.random.default_rng(14)
    data = w["data"].copy()
    bad = rng.choice(len(data), 150, replace=False)                          # 25 % gross outliers
    lo, hi = data.min(0) - 0.2, data.max(0) + 0.2
    data[bad] = (lo + rng.random((150, 3)) * (hi - lo)).astype(F)

    def err(R, t):
        ang = np.degrees(np.arccos(np.clip((np.trace(R @ w["R_true"].T) - 1) / 2, -1, 1)))
        return ang, float(np.linalg.norm(t - w["t_true"]))

    g = driver.FastGoICP(w["model"], data, 0.03, 1e-4, trim_fraction=0.3)
    R, t = g.run()
    ang, dt = err(R, t)
    mse = float(g.best_sse) / g.n_inliers
    g.close()
    assert g.n_inliers == 420 and ang < 2.0 and dt < 0.05 and mse < 5e-4
    g0 = driver.FastGoICP(w["model"], data, 0.03, 1e-4)
    R0, t0 = g0.run()
    ang0, dt0 = err(R0, t0)
    g0.close()
    assert float(g0.best_sse) / len(data) > 10 * mse                          # the untrimmed objective is dominated by the outliers
    assert ang <= ang0 + 0.5


@pytest.mark.gpu
def test_cuda_trimmed_bounds_beyond_the_shared_memory_size():
    """Round 1 capped trimmed bounds at 25,600 data points (2 x ns floats of shared memory per block).  Larger clouds keep
    the per-point terms in an L2-resident scratch slice per block: a 40,000-point data cloud against the definition --
    the per-point terms are the untrimmed bounds of one-point clouds... checked here through the identity
    trimmed(rho -> 0+) == untrimmed and through the oracle on a subsample-sized cloud built from the same points."""
    from fast_go_icp_b200 import capi, driver
    w = workloads.synthetic_pair(nt=20000, ns=40000 // 4, seed=17)
    rng = np.random.default_rng(3)
    data = np.concatenate([w["data"]] * 4) + rng.normal(scale=1e-3, size=(40000, 3)).astype(np.float32)
    pp = driver.preprocess(w["model"], data.astype(np.float32))
    res = 0.02
    ctx = capi.Context(pp["model"], pp["data"], pp["bbox_min"], pp["bbox_max"], res, flags=capi.BUILD_PACKED)
    try:
        ns = len(pp["data"])
        assert ns == 40000
        R, _ = O.rotation(0.1, -0.05, 0.2)
        tc = workloads.translation_cube_list(12, level=3, seed=4)
        lb0, ub0 = ctx.bounds_batch(R, 0.125, False, tc)                      # untrimmed
        # dropping ONE point (rho just above 1 / ns) removes exactly the largest per-point term of every cube
        K = ctx.set_trim(1.5 / ns)
        assert K == ns - 1
        lb1, ub1 = ctx.bounds_batch(R, 0.125, False, tc)
        assert np.all(ub1 <= ub0) and np.all(lb1 <= lb0) and np.any(ub1 < ub0)
        # against the oracle (CPU restatement with the same trimming rule), 1 ulp
        lut, dims = O.lut_build(pp["model"], pp["bbox_min"], pp["bbox_max"], res)
        for rho in (0.25, 1.5 / ns):
            K = ctx.set_trim(rho)
            with O.trimmed(K):
                lb, ub = ctx.bounds_batch(R, 0.125, False, tc)
                wl, wu = O.bounds(lut, dims, pp["bbox_min"], res, pp["data"], R, 0.125, False, tc)
                assert np.allclose(ub, wu, rtol=ULP, atol=0) and np.allclose(lb, wl, rtol=ULP, atol=0)
    finally:
        ctx.set_trim(0.0)
        ctx.close()

"""GPU tests at BASELINE.json's FULL sizes (W5: 100,000-point model, 10,000-point data, grid 269x368x357) through
size-independent properties -- the oracle cannot finish these sizes in seconds, so the checks are identities the
domain offers -- plus the edge cases of the C ABI (ragged cube lists, T not a multiple of the warp's cube group,
cubes far outside the grid, tiny clouds, argument errors).

Properties:
  * fix_rot upper bound == sum_i sample(R p_i + t)            (bounds vs the sampling entry point + numpy fp64 sum)
  * lb <= ub; a zero-span translation cube has lb == ub; rotation slack only lowers both
  * the result of a cube does not depend on its position in the list, on the list length, on the kernel
    (plain / phase-ordered) or on which other cubes share the launch                          (bit-exact)
  * sse(R, t) == sum of nn(R, t) distances; nn distances == brute-force mode distances        (bit-exact)
  * ICP restarted from its own result stops within two iterations, never worse, pose within 1e-2 (idempotence up to
    the slow drift the 0.05 % stop rule leaves)
  * every data point's nearest-neighbour distance is consistent with the distance grid: |sqrt(d2) - sqrt(T[n])|
    <= distance to the nearest grid node n                                                     (triangle inequality)
"""
import numpy as np
import pytest

from fast_go_icp_b200 import capi, driver, workloads

pytestmark = pytest.mark.gpu
F = np.float32


@pytest.fixture(scope="module")
def w5():
    w = workloads.synthetic_pair(nt=100_000, ns=10_000, sigma=0.01, seed=1234)
    pp = driver.preprocess(w["model"], w["data"])
    ctx = capi.Context(pp["model"], pp["data"], pp["bbox_min"], pp["bbox_max"], 0.005, flags=capi.BUILD_PACKED)
    yield w, pp, ctx
    ctx.close()


def _rotate(R9, p):
    R = np.asarray(R9, F).reshape(3, 3).T          # column-major -> math matrix
    return p @ R.T


def test_fixrot_upper_bound_is_the_sum_of_samples(w5):
    _, pp, ctx = w5
    rot = workloads.rotation_cube_list(3, seed=31)
    tc = workloads.translation_cube_list(8, level=3, seed=32)
    for r in range(3):
        R, _ = driver.rotation_matrix(*rot[r, :3])
        lb, ub = ctx.bounds_batch(R, float(rot[r, 3]), True, tc)
        # canonical per-point arithmetic of R p + t (fma association of the kernels), done in numpy float32
        p = pp["data"]
        q = np.empty_like(p)
        for a in range(3):
            prod = (F(R[3 + a]) * p[:, 1]).astype(F)
            acc = (np.float64(R[a]) * p[:, 0].astype(np.float64) + prod.astype(np.float64)).astype(F)     # fma
            q[:, a] = (np.float64(R[6 + a]) * p[:, 2].astype(np.float64) + acc.astype(np.float64)).astype(F)
        for c in range(len(tc)):
            d2 = ctx.lut_sample((q + tc[c, :3]).astype(F), capi.SAMPLER_PACKED)
            d = np.sqrt(d2.astype(F)).astype(F)
            want_ub = float(np.sum((d * d).astype(F).astype(np.float64)))
            e = np.maximum(d - F(np.float32(1.732050807568877) * tc[c, 3]), 0)
            assert abs(ub[c] - want_ub) <= 2e-6 * want_ub
            assert lb[c] <= ub[c] and abs(lb[c] - float(np.sum((e * e).astype(np.float64)))) <= 2e-5 * want_ub


def test_bound_invariants_full_size(w5):
    _, _, ctx = w5
    rot, tc = workloads.bound_microbench(64, 32, seed=41)
    tc[:, ::4, 3] = 0.0                                            # every fourth cube has zero span
    lb0, ub0 = ctx.bounds_multi(rot, True, tc)
    lb1, ub1 = ctx.bounds_multi(rot, False, tc)
    assert np.all(lb0 <= ub0) and np.all(lb1 <= ub1)
    assert np.array_equal(lb0[:, ::4], ub0[:, ::4]) and np.array_equal(lb1[:, ::4], ub1[:, ::4])
    assert np.all(ub1 <= ub0) and np.all(lb1 <= lb0)              # rotation slack only lowers the bounds
    assert np.all(np.isfinite(ub0)) and np.all(ub0 >= 0)


def test_cube_results_do_not_depend_on_the_launch(w5):
    _, _, ctx = w5
    rot, tc = workloads.bound_microbench(256, 32, seed=43)
    ctx.set_phased(True)
    lb, ub = ctx.bounds_multi(rot, False, tc)
    ctx.set_phased(False)
    lbp, ubp = ctx.bounds_multi(rot, False, tc)
    assert np.array_equal(lb, lbp) and np.array_equal(ub, ubp)     # phase-ordered == plain, 8,192 cubes x 10,000 points
    ctx.set_phased(True)
    # permutation of the rotation cubes and of each cube list
    rng = np.random.default_rng(44)
    pr = rng.permutation(256)
    pc = rng.permutation(32)
    lb2, ub2 = ctx.bounds_multi(rot[pr], False, np.ascontiguousarray(tc[pr][:, pc]))
    assert np.array_equal(lb2, lb[pr][:, pc]) and np.array_equal(ub2, ub[pr][:, pc])
    # a sub-list evaluated alone (different kernel geometry, plain kernel: < 4096 pairs)
    lb3, ub3 = ctx.bounds_multi(rot[:7], False, np.ascontiguousarray(tc[:7, :13]))
    assert np.array_equal(lb3, lb[:7, :13]) and np.array_equal(ub3, ub[:7, :13])
    # one rotation cube through the single-rotation entry point
    R, _ = driver.rotation_matrix(*rot[5, :3])
    lb4, ub4 = ctx.bounds_batch(R, float(rot[5, 3]), False, tc[5])
    assert np.array_equal(lb4, lb[5]) and np.array_equal(ub4, ub[5])


def test_ragged_and_odd_cube_lists(w5):
    """T that is not a multiple of 4 or 32, T > 32, unused slots (negative span) in the phase-ordered kernel."""
    _, _, ctx = w5
    rot = workloads.rotation_cube_list(200, seed=51)
    full = np.stack([workloads.translation_cube_list(70, level=4, seed=500 + r) for r in range(200)])
    ctx.set_phased(False)
    lbr, ubr = ctx.bounds_multi(rot, False, full)
    ctx.set_phased(True)
    for T in (1, 3, 33, 70):
        lb, ub = ctx.bounds_multi(rot, False, np.ascontiguousarray(full[:, :T]))
        assert np.array_equal(lb, lbr[:, :T]) and np.array_equal(ub, ubr[:, :T])
    ragged = full[:, :32].copy()
    rng = np.random.default_rng(52)
    unused = rng.random((200, 32)) < 0.4
    ragged[unused, 3] = -1.0                                       # unused slots
    lb, ub = ctx.bounds_multi(rot, False, ragged)                  # 6400 pairs -> phase-ordered kernel
    assert np.array_equal(lb[~unused], lbr[:, :32][~unused]) and np.array_equal(ub[~unused], ubr[:, :32][~unused])


def test_cubes_far_outside_the_grid_clamp_to_the_border(w5):
    _, pp, ctx = w5
    I = np.array([1, 0, 0, 0, 1, 0, 0, 0, 1], F)
    far = F([[50.0, 0, 0, 0.0], [0, -80.0, 3.0, 0.0], [1e4, 1e4, -1e4, 0.0]])
    lb, ub = ctx.bounds_batch(I, 0.0, True, far)
    assert np.all(np.isfinite(ub)) and np.array_equal(lb, ub)
    # clamp-to-edge: every query samples a border texel, so the bound cannot exceed ns * max(grid)
    lut, _ = ctx.lut_download()
    assert np.all(ub <= len(pp["data"]) * float(lut.max()) * (1 + 1e-6))
    q = (pp["data"][:100] + far[2, :3]).astype(F)
    assert np.array_equal(ctx.lut_sample(q, capi.SAMPLER_PACKED), ctx.lut_sample(q, capi.SAMPLER_GRID))


def test_nn_sse_consistency_full_size(w5):
    w, pp, ctx = w5
    lut, dims = ctx.lut_download()
    grid = lut.reshape(dims[2], dims[1], dims[0])
    for rotv, t in [((0, 0, 0), (0, 0, 0)), ((0.3, -0.2, 0.1), (0.1, -0.05, 0.02))]:
        R, _ = driver.rotation_matrix(*F(rotv))
        t = F(t)
        idx, d2 = ctx.nn(R, t, False)
        assert idx.min() >= 0 and idx.max() < len(pp["model"])
        sse = ctx.sse(R, t)
        assert sse == F(np.sum(d2.astype(np.float64)))             # fp64 sum of the per-point terms, rounded once
        ctx.set_nn_mode(1)
        idx_b, d2_b = ctx.nn(R, t, False)
        ctx.set_nn_mode(0)
        assert np.array_equal(idx, idx_b) and np.array_equal(d2, d2_b)   # cell-grid search == tiled brute force
        idx_r, d2_r = ctx.nn(R, t, True)
        assert np.mean(idx_r == idx) > 0.999                       # rooted and squared rules differ only on near-ties
        # triangle inequality against the distance grid at the nearest node
        q = _rotate(R, pp["data"]) + t
        u = (q - pp["bbox_min"]) / F(0.005)
        n = np.clip(np.rint(u).astype(int), 0, np.array(dims) - 1)
        node_d = np.sqrt(grid[n[:, 2], n[:, 1], n[:, 0]])
        off = np.linalg.norm(q - (pp["bbox_min"] + n * F(0.005)), axis=1)
        assert np.all(np.abs(np.sqrt(d2) - node_d) <= off * 1.001 + 1e-5)


def test_far_queries_through_the_coarse_boxes_equal_brute_force(w5):
    """Poses that throw most of the data cloud far from the model (the queries whose ball covers thousands of cell rows):
    the cell-grid search with the bounding-box culling of far queries (default), with the culling forced on for every
    query (2) and switched off (3) all return the brute-force winners, indices and distance bits, under both tie rules."""
    w, pp, ctx = w5
    for rotv, t in [((0.5, 0.5, -0.4), (0.9, -0.7, 0.8)), ((-0.2, 0.7, 0.3), (-1.0, 1.0, -1.0)), ((0.1, 0.0, 0.0), (0.3, 0.2, -0.25))]:
        R, _ = driver.rotation_matrix(*F(rotv))
        t = F(t)
        for rooted in (False, True):
            ctx.set_nn_mode(1)
            want = ctx.nn(R, t, rooted)
            for mode in (0, 2, 3):
                ctx.set_nn_mode(mode)
                got = ctx.nn(R, t, rooted)
                assert np.array_equal(got[0], want[0]) and np.array_equal(got[1], want[1]), (rotv, rooted, mode)
            ctx.set_nn_mode(0)
        assert np.sqrt(want[1]).mean() > 0.1 or rotv == (0.1, 0.0, 0.0)       # the first two poses really are far


def test_icp_is_idempotent_at_its_fixed_point(w5):
    _, _, ctx = w5
    I = np.array([1, 0, 0, 0, 1, 0, 0, 0, 1], F)
    e, R, t, it = ctx.icp(I, np.zeros(3, F), 100, 0.0005)
    e2, R2, t2, it2 = ctx.icp(R, t, 100, 0.0005)
    assert it2 <= 2 and abs(e2 - e) <= 1e-3 * e
    assert np.allclose(R2, R, atol=1e-2) and np.allclose(t2, t, atol=1e-2)     # the 0.05 % stop rule leaves a slow drift
    assert e2 <= e * (1 + 1e-6)


def test_morton_storage_order_changes_no_result():
    """The library keeps the data cloud in Morton order on the device.  Every result must equal the one
    obtained with the caller's order kept (FGOICP_BUILD_KEEP_ORDER): bounds, SSE, per-point NN outputs (returned in the
    caller's order), ICP, inner searches, trimmed variants included."""
    w = workloads.synthetic_pair(nt=20_000, ns=3_000, sigma=0.01, seed=21)
    pp = driver.preprocess(w["model"], w["data"])
    a = capi.Context(pp["model"], pp["data"], pp["bbox_min"], pp["bbox_max"], 0.01, flags=capi.BUILD_PACKED)
    b = capi.Context(pp["model"], pp["data"], pp["bbox_min"], pp["bbox_max"], 0.01,
                     flags=capi.BUILD_PACKED | capi.BUILD_KEEP_ORDER)
    try:
        rot, tc = workloads.bound_microbench(160, 32, seed=23)
        R, _ = driver.rotation_matrix(F(0.2), F(-0.1), F(0.15))
        t = F([0.03, -0.02, 0.05])
        for rho in (0.0, 0.2):
            a.set_trim(rho); b.set_trim(rho)
            for fix_rot in (True, False):
                la, ua = a.bounds_multi(rot, fix_rot, tc)
                lb_, ub_ = b.bounds_multi(rot, fix_rot, tc)
                assert np.array_equal(la, lb_) and np.array_equal(ua, ub_)
            assert a.sse(R, t) == b.sse(R, t)
            ea, Ra, ta, ia = a.icp(R, t, 100, 0.005)
            eb, Rb, tb, ib = b.icp(R, t, 100, 0.005)
            assert ia == ib and ea == eb and np.array_equal(Ra, Rb) and np.array_equal(ta, tb)
            thr = a.ns * 1e-4
            ua, ta_, eva = a.bnb_r3_batch(rot[:6], True, 1e10, thr)
            ub2, tb_, evb = b.bnb_r3_batch(rot[:6], True, 1e10, thr)
            assert np.array_equal(eva, evb) and np.array_equal(ua, ub2) and np.array_equal(ta_, tb_)
        a.set_trim(0.0); b.set_trim(0.0)
        for rooted in (False, True):
            ia_, da = a.nn(R, t, rooted)
            ib_, db = b.nn(R, t, rooted)
            assert np.array_equal(ia_, ib_) and np.array_equal(da, db)
    finally:
        a.close(); b.close()


def test_tiny_clouds_and_argument_errors():
    model = F([[0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1]])
    data = F([[0.1, 0.1, 0.1]])
    ctx = capi.Context(model, data, model.min(0), model.max(0), 0.25, flags=capi.BUILD_PACKED)
    I = np.array([1, 0, 0, 0, 1, 0, 0, 0, 1], F)
    idx, d2 = ctx.nn(I, np.zeros(3, F), False)
    assert idx[0] == 0 and np.isclose(d2[0], 0.03, rtol=1e-6)
    lb, ub = ctx.bounds_batch(I, 0.0, True, F([[0, 0, 0, 0.0]]))
    assert lb[0] == ub[0] and ub[0] >= 0
    e, R, t, it = ctx.icp(I, np.zeros(3, F), 5, 0.05)
    assert np.isfinite(e)
    ub1, bt1, ev1 = ctx.bnb_r3(F([0, 0, 0, 0.5]), True, 1e10, 1e-3)
    assert np.isfinite(ub1) and ev1 >= 1
    with pytest.raises(capi.FgoicpError):
        ctx.bounds_batch(I, 0.0, True, np.zeros((0, 4), F))                          # T = 0
    with pytest.raises(capi.FgoicpError):
        ctx.set_sampler(capi.SAMPLER_TEX)                                            # texture not built
    ctx.close()
    with pytest.raises(capi.FgoicpError):
        capi.Context(model, data, model.min(0), model.max(0), 1e-4, flags=0)          # grid dims >= 2048
    with pytest.raises(capi.FgoicpError):
        capi.Context(model[:0], data, model.min(0), model.max(0), 0.25, flags=0)      # empty model

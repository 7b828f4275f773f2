"""GPU parity tests: every C-ABI entry point against the CPU oracle on the same seeded inputs, and --
when oracle/_ref/libfgoicp_ref.so is present -- against the unmodified reference kernels.

Tolerances (stated per test):
  * grid cells, NN indices, per-point squared distances: bit-exact;
  * per-cube bounds and SSE: sums are fp64-accumulated on both sides and rounded once, so they agree
    to 1 ulp of fp32 (rtol 2.4e-7), far inside BASELINE.json's "stated fp32 tolerance";
  * against the reference's texture unit and CUB reduction order: rtol 2e-4 on bounds.
"""
import numpy as np
import pytest

from fast_go_icp_b200 import capi, workloads
from oracle import oracle as O
from oracle import ref as REF

pytestmark = pytest.mark.gpu
ULP = 2.4e-7


def _lut(pp):
    return pp["lut"], pp["dims"], pp["bbox_min"], float(pp["res"])


def test_grid_build_is_bit_exact(small_problem, gpu_ctx):
    pp = small_problem
    got, dims = gpu_ctx.lut_download()
    assert np.array_equal(dims, pp["dims"])
    assert np.array_equal(got, pp["lut"])           # hierarchical build == oracle brute force, every cell
    brute = capi.Context(pp["model"], pp["data"], pp["bbox_min"], pp["bbox_max"], float(pp["res"]),
                         flags=capi.BUILD_BRUTE_LUT)
    assert np.array_equal(brute.lut_download()[0], got)
    brute.close()


def test_samplers(small_problem, gpu_ctx):
    pp = small_problem
    rng = np.random.default_rng(10)
    q = rng.uniform(-1.4, 1.4, (20000, 3)).astype(np.float32)
    want = O.lut_sample(*_lut(pp), q)
    g = gpu_ctx.lut_sample(q, capi.SAMPLER_GRID)
    p = gpu_ctx.lut_sample(q, capi.SAMPLER_PACKED)
    t = gpu_ctx.lut_sample(q, capi.SAMPLER_TEX)
    assert np.array_equal(g, want)                  # manual filter: bit-exact with the oracle
    assert np.array_equal(p, g)                     # corner-packed layout: same arithmetic, same bits
    # hardware filter: same semantics (half-texel shift, clamp, 8-bit weights), but the texture unit's
    # internal arithmetic is not IEEE fp32: measured on B200 (profiles/tex_conformance_r01.json) the
    # median relative difference is ~2e-6 and rare samples next to the surface differ by up to ~3 %
    scale = np.maximum(np.abs(want), 1e-6)
    rel = np.abs(t - want) / scale
    assert np.median(rel) < 1e-5 and np.quantile(rel, 0.99) < 5e-3 and np.max(rel) < 0.1


def test_device_sin_equals_the_reference_builds_constants(gpu_ctx):
    """sin(span * sqrt3 * pi / 2) (registration.cu:41-42): the library's device values equal, bit for bit, the constants
    a kernel compiled inside the reference build produced (tests/golden/reference_sin.json, what the oracle uses) -- and,
    when oracle/_ref is present, what that kernel returns on this very GPU."""
    from conftest import install_reference_sin
    spans, want = install_reference_sin()
    dev = gpu_ctx.rot_sin(spans)
    assert np.array_equal(dev.view(np.uint32), want.view(np.uint32))
    if REF.available():
        assert np.array_equal(REF.rot_sin(spans).view(np.uint32), want.view(np.uint32))


@pytest.mark.parametrize("fix_rot", [True, False])
@pytest.mark.parametrize("sampler", [capi.SAMPLER_GRID, capi.SAMPLER_PACKED])
def test_bounds_batch_vs_oracle(small_problem, gpu_ctx, fix_rot, sampler):
    pp = small_problem
    gpu_ctx.set_sampler(sampler)
    for k, (rot, span) in enumerate([((0.1, -0.2, 0.05), 0.125), ((0.5, 0.5, -0.5), 0.5), ((0, 0, 0), 0.0625)]):
        R, _ = O.rotation(*np.float32(rot))
        for T in (1, 7, 32, 45):
            tc = workloads.translation_cube_list(T, level=1 + (k + T) % 4, seed=T)
            lb, ub = gpu_ctx.bounds_batch(R, span, fix_rot, tc)
            wl, wu = O.bounds(*_lut(pp), pp["data"], R, span, fix_rot, tc)
            assert np.allclose(ub, wu, rtol=ULP, atol=0) and np.allclose(lb, wl, rtol=ULP, atol=0)
    gpu_ctx.set_sampler(capi.SAMPLER_PACKED)


def test_bounds_multi_matches_batch_and_is_deterministic(small_problem, gpu_ctx):
    rot = workloads.rotation_cube_list(40, seed=3)
    rot[:10, 3] = 0.25
    _, tc = workloads.bound_microbench(40, 32, seed=5)
    for fix_rot in (True, False):
        lb, ub = gpu_ctx.bounds_multi(rot, fix_rot, tc)
        lb2, ub2 = gpu_ctx.bounds_multi(rot, fix_rot, tc)
        assert np.array_equal(lb, lb2) and np.array_equal(ub, ub2)
        for r in (0, 9, 39):
            R, _ = O.rotation(*rot[r, :3])
            l1, u1 = gpu_ctx.bounds_batch(R, float(rot[r, 3]), fix_rot, tc[r])
            assert np.allclose(l1, lb[r], rtol=ULP) and np.allclose(u1, ub[r], rtol=ULP)


def test_phase_ordered_kernel_equals_plain_kernel(small_problem, gpu_ctx):
    """The z-phase-ordered evaluation reorders the same per-point terms; sums are fp64, so floats are equal."""
    rng = np.random.default_rng(31)
    for n_rot, T in ((200, 32), (137, 45), (300, 20)):
        rot = workloads.rotation_cube_list(n_rot, seed=n_rot)
        rot[::7, 3] = 0.25
        tc = np.stack([workloads.translation_cube_list(T, level=2 + (r % 3), seed=r) for r in range(n_rot)])
        tc[rng.random((n_rot, T)) < 0.2, 3] = 0.0          # zero-span cubes are legal
        for fix_rot in (True, False):
            gpu_ctx.set_phased(False)
            lb0, ub0 = gpu_ctx.bounds_multi(rot, fix_rot, tc)
            gpu_ctx.set_phased(True)
            lb1, ub1 = gpu_ctx.bounds_multi(rot, fix_rot, tc)
            assert np.array_equal(lb0, lb1) and np.array_equal(ub0, ub1)
    gpu_ctx.set_phased(True)


def test_bounds_multi_dev_and_best_ub(small_problem, gpu_ctx):
    import torch
    rot = workloads.rotation_cube_list(300, seed=8)
    _, tc = workloads.bound_microbench(300, 32, seed=9)
    lb, ub = gpu_ctx.bounds_multi(rot, True, tc)
    d_rot, d_tc = torch.from_numpy(rot).cuda(), torch.from_numpy(tc).cuda()
    d_lb, d_ub = torch.empty(300, 32, device="cuda"), torch.empty(300, 32, device="cuda")
    d_best = torch.empty(1, device="cuda")
    stream = torch.cuda.Stream()
    stream.wait_stream(torch.cuda.current_stream())
    gpu_ctx.set_stream(stream.cuda_stream)
    gpu_ctx.bounds_multi_dev(d_rot.data_ptr(), 300, True, d_tc.data_ptr(), 32, d_lb.data_ptr(), d_ub.data_ptr(),
                             d_best.data_ptr())
    torch.cuda.synchronize()
    gpu_ctx.set_stream(0)
    assert np.array_equal(d_lb.cpu().numpy(), lb) and np.array_equal(d_ub.cpu().numpy(), ub)
    assert float(d_best.item()) == float(ub.min())


@pytest.mark.parametrize("nn_mode", [0, 1, 2, 3])
def test_nn_indices_bit_exact(small_problem, gpu_ctx, nn_mode):
    pp = small_problem
    gpu_ctx.set_nn_mode(nn_mode)
    for rot, t in [((0.2, 0.1, -0.1), (0.05, -0.02, 0.01)), ((0, 0, 0), (0, 0, 0)), ((-0.4, 0.3, 0.2), (0.4, 0.3, -0.5))]:
        R, _ = O.rotation(*np.float32(rot))
        t = np.float32(t)
        for rooted in (False, True):
            idx, d2 = gpu_ctx.nn(R, t, rooted)
            widx, wd2 = O.nn(pp["model"], pp["data"], R, t, rooted)
            assert np.array_equal(idx, widx)
            assert np.array_equal(d2, wd2)
        sse = gpu_ctx.sse(R, t)
        assert sse == O.sse(pp["model"], pp["data"], R, t)
    gpu_ctx.set_nn_mode(0)


def test_nn_duplicate_points_lowest_index_wins():
    rng = np.random.default_rng(12)
    base = rng.uniform(-0.8, 0.8, (700, 3)).astype(np.float32)
    model = np.concatenate([base, base[::-1], base])          # every point three times
    data = base[:257] + np.float32(1e-3)
    ctx = capi.Context(model, data, model.min(0), model.max(0), 0.1, flags=0)
    I = np.eye(3, dtype=np.float32).ravel()
    for nn_mode in (0, 1, 2, 3):
        ctx.set_nn_mode(nn_mode)
        for rooted in (False, True):
            for t in (np.zeros(3, np.float32), np.float32([0.7, -0.4, 1.3])):      # near and far queries
                idx, _ = ctx.nn(I, t, rooted)
                widx, _ = O.nn(model, data, I, t, rooted)
                assert np.array_equal(idx, widx) and np.all(idx < 700)
    ctx.close()


def test_nn_rooted_rule_near_ties():
    """Distinct squared distances that round to the SAME square root: the rooted rule (icp3d.cu:17-26) keeps the
    first such index, the squared rule (registration.cu:160-172) the strictly smaller d2.  The CUDA search finds
    the squared winner first and must fall back to the exact rooted scan exactly in this window."""
    rng = np.random.default_rng(21)
    filler = rng.uniform(-0.9, 0.9, (500, 3)).astype(np.float32)
    filler = filler[np.linalg.norm(filler, axis=1) > 0.6]        # keep the neighbourhood of the probes empty
    data = np.zeros((64, 3), np.float32)
    model = [np.float32([-0.95, -0.95, -0.95]), np.float32([0.95, 0.95, 0.95])]
    # probe k: query q_k = (0.01 k, 0, 0); model gets p_a = q + (r, 0, dz) then p_b = q + (r, 0, 0) with
    # d2(p_a) one or two ulps above d2(p_b) = r*r
    for k in range(64):
        q = np.float32([0.004 * k - 0.128, 0.0, 0.0])
        data[k] = q
    # the probes sit on a ring of radius 0.25 around their query, far from each other's queries
    want_cases = 0
    for k in range(64):
        q = data[k]
        r = np.float32(0.25)
        dz = np.float32(r * np.sqrt(np.float32(2.0 ** -23)) * (1.0 + 0.1 * (k % 5)))
        model.append((q + np.float32([0, r, dz])).astype(np.float32))     # earlier index, slightly farther
        model.append((q + np.float32([0, r, 0])).astype(np.float32))      # later index, nearest in d2
    model = np.concatenate([np.array(model, np.float32), filler]).astype(np.float32)
    ctx = capi.Context(model, data, model.min(0), model.max(0), 0.05, flags=0)
    I = np.eye(3, dtype=np.float32).ravel()
    z = np.zeros(3, np.float32)
    for nn_mode in (0, 1, 2, 3):
        ctx.set_nn_mode(nn_mode)
        idx_r, d_r = ctx.nn(I, z, True)
        idx_s, d_s = ctx.nn(I, z, False)
        w_r, wd_r = O.nn(model, data, I, z, True)
        w_s, wd_s = O.nn(model, data, I, z, False)
        assert np.array_equal(idx_r, w_r) and np.array_equal(idx_s, w_s)
        assert np.array_equal(d_r, wd_r) and np.array_equal(d_s, wd_s)
        want_cases = int(np.sum(w_r != w_s))
    assert want_cases >= 8, "the construction must actually produce rooted/squared disagreements (%d)" % want_cases
    ctx.close()


def test_icp_vs_oracle(small_problem, gpu_ctx):
    pp = small_problem
    I = np.eye(3, dtype=np.float32).ravel()
    for R0, t0, thr in [(I, np.zeros(3, np.float32), 0.05), (O.rotation(0.3, 0.1, -0.2)[0], np.float32([0.1, 0, -0.1]), 0.005),
                        (O.rotation(-0.1, 0.05, 0.1)[0], np.float32([0.02, 0.03, 0.0]), 0.0005)]:
        e, R, t, it = gpu_ctx.icp(R0, t0, 100, thr)
        we, wR, wt, wit = O.icp(pp["model"], pp["data"], 100, thr, R0, t0)
        assert it == wit
        assert abs(e - we) <= 1e-6 * we                        # BASELINE.json: MSE within 1e-6 relative
        assert np.allclose(R, wR, atol=2e-6) and np.allclose(t, wt, atol=2e-6)
    e, R, t, it = gpu_ctx.icp(I, np.zeros(3, np.float32), 0, 0.05)   # max_iter = 0: returns the seed state
    assert it == 0


@pytest.mark.parametrize("fix_rot", [True, False])
def test_bnb_r3_vs_oracle(small_problem, gpu_ctx, fix_rot):
    pp = small_problem
    thr = len(pp["data"]) * 1e-4
    cubes = np.float32([[0.25, -0.25, 0.25, 0.25], [0.0625, 0.1875, -0.0625, 0.0625], [-0.5, 0.5, 0.5, 0.5],
                        [0.125, 0.125, 0.125, 0.125], [-0.1875, 0.0625, 0.3125, 0.0625]])
    for best_sse in (1e10, 5.0, 0.5):
        ub, bt, ev = gpu_ctx.bnb_r3_batch(cubes, fix_rot, best_sse, thr)
        for i, c in enumerate(cubes):
            wub, wbt, wev, _ = O.bnb_r3(pp["model"], pp["data"], *_lut(pp), c, fix_rot, best_sse, thr)
            assert ev[i] == wev, (i, best_sse)                 # same traversal: same number of evaluations
            assert np.isclose(ub[i], wub, rtol=ULP, atol=0) and np.array_equal(bt[i], wbt)
        u1, t1, e1 = gpu_ctx.bnb_r3(cubes[1], fix_rot, best_sse, thr)
        assert u1 == ub[1] and np.array_equal(t1, bt[1]) and e1 == ev[1]


@pytest.mark.parametrize("fix_rot", [True, False])
def test_bnb_round_synchronous_equals_persistent(small_problem, gpu_ctx, fix_rot):
    """The round-synchronous schedule (all searches of a level advance together, bounds of every round through
    the phase-ordered kernel or -- for the last stragglers -- the plain one) must take exactly the decisions of
    the persistent per-cube kernel: same best_ub bits, same best_t, same number of evaluations, for every cube;
    and both must match the oracle's traversal."""
    pp = small_problem
    thr = len(pp["data"]) * 1e-4
    cubes = workloads.rotation_cube_list(160, seed=5)
    cubes[:40, 3] = 0.125                                    # mixed spans: different rotation slack per cube
    import os
    for best_sse, min_pairs in ((1e10, "0"), (3.0, "64"), (0.6, "100000")):
        os.environ["FGOICP_BNBR_MIN_PAIRS"] = min_pairs      # phased for every round / mixed / plain kernel only
        gpu_ctx.set_bnb_mode(1)
        ub1, bt1, ev1 = gpu_ctx.bnb_r3_batch(cubes, fix_rot, best_sse, thr)
        gpu_ctx.set_bnb_mode(2)
        ub2, bt2, ev2 = gpu_ctx.bnb_r3_batch(cubes, fix_rot, best_sse, thr)
        gpu_ctx.set_bnb_mode(0)
        assert np.array_equal(ev1, ev2)
        assert np.array_equal(ub1.view(np.uint32), ub2.view(np.uint32))
        assert np.array_equal(bt1, bt2)
        for i in (0, 57, 159):
            wub, wbt, wev, _ = O.bnb_r3(pp["model"], pp["data"], *_lut(pp), cubes[i], fix_rot, best_sse, thr)
            assert ev2[i] == wev and np.isclose(ub2[i], wub, rtol=ULP, atol=0) and np.array_equal(bt2[i], wbt)
    os.environ.pop("FGOICP_BNBR_MIN_PAIRS", None)


@pytest.mark.skipif(not REF.available(), reason="oracle/_ref not built")
def test_against_unmodified_reference_kernels(small_problem, gpu_ctx):
    """The reference's own CUDA code (real tex3D, per-cube launches, thrust reductions) on the same clouds."""
    pp = small_problem
    raw = pp["raw"]
    ref = REF.Reference(raw["model"], raw["data"], float(pp["res"]), 1e-4)
    rp = ref.preprocessed()
    for k in ("model", "data", "offset_pcs", "offset_pct", "bbox_min", "bbox_max"):
        assert np.array_equal(rp[k], pp[k]), k                 # preprocessing: bit-exact
    assert rp["scale"] == pp["scale"]
    rlut, rdims = ref.lut()
    assert np.array_equal(rdims, pp["dims"])
    assert np.array_equal(rlut, pp["lut"])                     # every LUT cell bit-exact vs the reference kernel
    rng = np.random.default_rng(21)
    q = rng.uniform(-1.2, 1.2, (20000, 3)).astype(np.float32)
    assert np.array_equal(ref.lut_sample(q), gpu_ctx.lut_sample(q, capi.SAMPLER_TEX))   # same hardware path
    # bounds: reference (texture + CUB fp32 tree) vs ours (manual filter + fp64 sums)
    gpu_ctx.set_sampler(capi.SAMPLER_PACKED)
    for rot in ([0.25, -0.25, 0.25, 0.25], [0.0625, 0.1875, -0.0625, 0.0625]):
        rot = np.float32(rot)
        R, _ = O.rotation(*rot[:3])
        tc = workloads.translation_cube_list(32, level=3, seed=4)
        for fix_rot in (True, False):
            rl, ru = ref.bounds(rot, fix_rot, tc)
            l, u = gpu_ctx.bounds_batch(R, float(rot[3]), fix_rot, tc)
            assert np.allclose(u, ru, rtol=1e-4, atol=2e-4) and np.allclose(l, rl, rtol=1e-4, atol=2e-4)
    # exact SSE and ICP
    R, _ = O.rotation(0.2, 0.1, -0.1)
    t = np.float32([0.05, -0.02, 0.01])
    assert abs(ref.sse(R, t) - gpu_ctx.sse(R, t)) <= 2e-6 * gpu_ctx.sse(R, t)
    I = np.eye(3, dtype=np.float32).ravel()
    re_, rR, rt = ref.icp(I, np.zeros(3, np.float32), 100, 0.05)
    e, R1, t1, _ = gpu_ctx.icp(I, np.zeros(3, np.float32), 100, 0.05)
    assert abs(re_ - e) <= 1e-5 * e and np.allclose(rR, R1, atol=1e-5) and np.allclose(rt, t1, atol=1e-5)
    ref.close()


# ---- device-side constructor preprocessing (fgoicp.cpp:176-287) ------------------------------

def _raw_clouds(nt, ns, seed):
    rng = np.random.default_rng(seed)
    model = (rng.normal(size=(nt, 3)) * [40, 25, 60] + [7, -300, 1e3]).astype(np.float32)
    data = (rng.normal(size=(ns, 3)) * [35, 20, 50] + [-3, 12, -90]).astype(np.float32)
    return model, data


@pytest.mark.parametrize("nt,ns", [(1, 2), (3, 2), (1023, 1024), (1025, 2049), (5000, 700), (110_000, 10_000),
                                   (1_000_003, 250_001)])
def test_device_preprocess_is_bit_exact(nt, ns):
    """Reference mode: centred + scaled clouds, both offsets, the scale and the target's range carry the SAME BITS as
    the oracle's serial fp32 restatement (centroid = sequential sum in index order).  Sizes straddle the kernel's
    1024-point tiles."""
    model, data = _raw_clouds(nt, ns, nt + ns)
    want = O.preprocess(model, data)
    got = capi.preprocess(model, data)
    for k in ("model", "data", "offset_pcs", "offset_pct", "bbox_min", "bbox_max"):
        assert np.array_equal(got[k], want[k]), k
    assert np.float32(got["scale"]) == np.float32(want["scale"])
    assert got["device_ms"] > 0


def test_device_preprocess_on_device_buffers():
    import torch
    model, data = _raw_clouds(20_000, 3_000, 3)
    want = O.preprocess(model, data)
    dm, dd = torch.from_numpy(model).cuda(), torch.from_numpy(data).cuda()
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        got = capi.preprocess_dev(dm.data_ptr(), len(model), dd.data_ptr(), len(data), cuda_stream_ptr=st.cuda_stream)
    assert np.array_equal(dm.cpu().numpy(), want["model"]) and np.array_equal(dd.cpu().numpy(), want["data"])
    assert np.array_equal(got["offset_pcs"], want["offset_pcs"]) and np.array_equal(got["bbox_max"], want["bbox_max"])


def test_device_preprocess_options():
    """TREE_CENTROID: deterministic fp64 reduction, within 1 fp32 ulp of the exact mean (NOT the reference's bits);
    SCALE_BOTH: both clouds inside [-1, 1]^3 (the translation domain, fgoicp.cpp:113)."""
    model, data = _raw_clouds(300_000, 40_000, 9)
    a = capi.preprocess(model, data, flags=capi.PRE_TREE_CENTROID)
    b = capi.preprocess(model, data, flags=capi.PRE_TREE_CENTROID)
    for k in ("model", "data", "offset_pcs", "offset_pct"):
        assert np.array_equal(a[k], b[k])                                       # repeatable bit for bit
    for k, cloud in (("offset_pcs", data), ("offset_pct", model)):
        exact = -cloud.astype(np.float64).mean(0)
        assert np.all(np.abs(a[k] - exact) <= np.spacing(np.abs(exact).astype(np.float32)))
    ref = O.preprocess(model, data)
    assert np.max(np.abs(ref["model"])) > 1.0                                   # the reference lets the target stick out
    c = capi.preprocess(model, data, flags=capi.PRE_SCALE_BOTH)
    assert np.max(np.abs(c["model"])) <= 1.0 and np.max(np.abs(c["data"])) <= 1.0
    assert max(np.max(np.abs(c["model"])), np.max(np.abs(c["data"]))) == 1.0
    assert np.array_equal(c["offset_pct"], ref["offset_pct"])                   # centring unchanged


def test_run_with_device_preprocess_is_identical():
    from fast_go_icp_b200 import driver
    w = workloads.synthetic_pair(nt=4000, ns=500, sigma=0.005, seed=8, max_angle=0.8)
    out = []
    for dev in (False, True):
        g = driver.FastGoICP(w["model"], w["data"], 0.02, 1e-4, device_preprocess=dev)
        R, t = g.run()
        out.append((R, t, g.best_sse))
        g.close()
    assert np.array_equal(out[0][0], out[1][0]) and np.array_equal(out[0][1], out[1][1]) and out[0][2] == out[1][2]


# ---- batched refinements on a slot pool with a device-side job queue -------------------------------------------

def test_icp_batch_equals_single_refinements(small_problem, gpu_ctx, monkeypatch):
    """23 seeds on 4 (then 64) instance slots: every slot is refilled several times on the device; each result must
    carry the same bits as the refinement run alone (IterativeClosestPoint3D.run(), icp3d.cu:55-108)."""
    rng = np.random.default_rng(31)
    seeds_R, seeds_t = [], []
    for k in range(23):
        v = rng.uniform(-0.45, 0.45, 3) * (0.1 if k % 3 == 0 else 1.0)      # easy and hard starts -> very different iteration counts
        R, _ = O.rotation(*v.astype(np.float32))
        seeds_R.append(R)
        seeds_t.append(rng.uniform(-0.2, 0.2, 3).astype(np.float32))
    single = [gpu_ctx.icp(R, t, 100, 0.005) for R, t in zip(seeds_R, seeds_t)]
    assert len({s[3] for s in single}) > 3                                    # iteration counts do differ
    for slots in ("4", "1", "64"):
        monkeypatch.setenv("FGOICP_ICP_SLOTS", slots)
        e, R, t, it = gpu_ctx.icp_batch(np.array(seeds_R), np.array(seeds_t), 100, 0.005)
        for k, (se, sR, st, sit) in enumerate(single):
            assert it[k] == sit and e[k] == np.float32(se), (slots, k)
            assert np.array_equal(R[k], sR) and np.array_equal(t[k], st), (slots, k)
    monkeypatch.delenv("FGOICP_ICP_SLOTS")
    # max_iter = 0 ends at the first loop head: seed pose, zero iterations (icp3d.cu:94, 106-107)
    e, R, t, it = gpu_ctx.icp_batch(np.array(seeds_R[:5]), np.array(seeds_t[:5]), 0, 0.005)
    assert np.all(it == 0) and np.array_equal(R, np.array(seeds_R[:5]).reshape(5, 9)) and np.all(e == np.float32(1e10))
    e0 = gpu_ctx.icp_batch(np.zeros((0, 9), np.float32), np.zeros((0, 3), np.float32), 10, 0.005)[0]
    assert len(e0) == 0


def test_persistent_icp_loop_equals_the_launch_chain(small_problem, gpu_ctx):
    """The persistent cooperative loop kernel (default) and the one-launch-per-stage chain walk the same ICP
    (icp3d.cu:85-108): same iteration counts, SSE, poses -- bit for bit -- for easy and hard starts, alone and batched."""
    rng = np.random.default_rng(77)
    seeds_R, seeds_t = [], []
    for k in range(17):
        v = rng.uniform(-0.45, 0.45, 3) * (0.1 if k % 2 == 0 else 1.0)
        seeds_R.append(O.rotation(*v.astype(np.float32))[0])
        seeds_t.append(rng.uniform(-0.2, 0.2, 3).astype(np.float32))
    out = {}
    try:
        for mode in (2, 1, 0):
            gpu_ctx.set_icp_mode(mode)
            out[mode] = [gpu_ctx.icp_batch(np.array(seeds_R), np.array(seeds_t), 100, thr) for thr in (0.05, 0.005, 0.0005)]
            out[mode].append(tuple(np.asarray(x) for x in gpu_ctx.icp(seeds_R[3], seeds_t[3], 100, 0.005)))
    finally:
        gpu_ctx.set_icp_mode(0)
    for other in (1, 0):
        for a, b in zip(out[2], out[other]):
            for x, y in zip(a, b):
                assert np.array_equal(x, y)
    assert out[2][1][3].max() > 10


@pytest.mark.parametrize("which", ["synthetic", "dragon"])
def test_scan_schedules_of_the_icp_loop_change_no_result(which):
    """The persistent ICP kernel scans a heavy query with four warps together and everything else with one warp each.
    Forcing EVERY scan to be cooperative (FGOICP_NN_HEAVY_ROWS=1), none (=0), the launch-chain driver
    (FGOICP_ICP_MODE=1) and the bounding-box culling switched off (FGOICP_NN_COARSE=0) must give the same SSE, poses and
    iteration counts as the default, bit for bit."""
    import json
    import os
    import subprocess
    import sys
    probe = os.path.join(os.path.dirname(os.path.abspath(__file__)), "icp_probe.py")
    outs = {}
    for tag, env in (("default", {}), ("all-coop", {"FGOICP_NN_HEAVY_ROWS": "1"}), ("no-coop", {"FGOICP_NN_HEAVY_ROWS": "0"}),
                     ("chain", {"FGOICP_ICP_MODE": "1"}), ("loop", {"FGOICP_ICP_MODE": "2"}), ("loop-all-coop", {"FGOICP_ICP_MODE": "2", "FGOICP_NN_HEAVY_ROWS": "1"}),
                     ("no-boxes", {"FGOICP_NN_COARSE": "0"})):
        r = subprocess.run([sys.executable, probe, which], capture_output=True, text=True, env=dict(os.environ, **env), timeout=300)
        assert r.returncode == 0, (tag, r.stderr[-2000:])
        outs[tag] = json.loads([l for l in r.stdout.splitlines() if l.startswith("PROBE ")][-1][6:])
    for tag in outs:
        assert outs[tag] == outs["default"], tag


@pytest.mark.parametrize("which", ["synthetic", "dragon"])
def test_winner_memo_changes_no_result(which):
    """The winner memo of the ICP searches (skip the scan when the point moved less than the clearance the last scan
    proved) must not change a single bit: refinements with the memo off (margin 0), at the default margin and at a
    large one give identical SSE, poses and iteration counts."""
    import json
    import os
    import subprocess
    import sys
    probe = os.path.join(os.path.dirname(os.path.abspath(__file__)), "icp_probe.py")
    outs = {}
    for margin in ("0", "2.5e-4", "4e-3"):
        env = dict(os.environ, FGOICP_NN_MARGIN=margin)
        r = subprocess.run([sys.executable, probe, which], capture_output=True, text=True, env=env, timeout=300)
        assert r.returncode == 0, r.stderr[-2000:]
        outs[margin] = json.loads([l for l in r.stdout.splitlines() if l.startswith("PROBE ")][-1][6:])
    assert outs["0"] == outs["2.5e-4"] == outs["4e-3"]
    its = outs["0"]["1e-05"]["it"]
    assert max(its) > 30                                  # long runs were exercised

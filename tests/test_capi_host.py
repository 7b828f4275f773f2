"""CPU tests: the C-ABI library loads, exports every symbol the header declares, fails loudly without a
GPU, and the host-side driver arithmetic matches the oracle bit for bit."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from fast_go_icp_b200 import capi, driver
from oracle import oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_functions():
    src = open(os.path.join(ROOT, "include", "fgoicp_c.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(fgoicp_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    L = capi.lib()
    names = _declared_functions()
    assert len(names) >= 20
    for n in names:
        assert hasattr(L, n), "missing export: " + n
    assert set(names) == set(capi.EXPORTS)
    assert b"sm_100a" in L.fgoicp_version()


def test_header_is_plain_c_and_struct_layouts_match_the_binding(tmp_path):
    """include/fgoicp_c.h compiles as strict C99, and the ctypes mirrors in capi.py have the compiler's layout."""
    import subprocess
    exe = str(tmp_path / "abi_sizes")
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I" + os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "tests", "cpp", "abi_sizes.c"), "-o", exe], check=True)
    out = dict((l.split()[0], [int(x) for x in l.split()[1:]]) for l in subprocess.run([exe], capture_output=True, text=True).stdout.splitlines())
    I, S, N = capi.Info, capi.LevelStats, capi.Normalisation
    assert out["fgoicp_info"] == [ctypes.sizeof(I), I.dims.offset, I.grid_bytes.offset, I.build_ms.offset]
    assert out["fgoicp_level_stats"] == [ctypes.sizeof(S), S.n_icp.offset, S.ms_bnb_ub.offset, S.best_icp_index.offset]
    assert out["fgoicp_normalisation"] == [ctypes.sizeof(N), N.scale.offset, N.bbox_min.offset, N.device_ms.offset]


def test_cpp_api_symbols_present():
    out = os.popen("nm -DC %s" % capi.LIB_PATH).read()
    assert "icp::FastGoICP::run()" in out
    assert "icp::FastGoICP::FastGoICP(" in out


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback_without_gpu():
    pts = np.random.default_rng(0).uniform(-1, 1, (16, 3)).astype(np.float32)
    with pytest.raises(capi.FgoicpError) as e:
        capi.Context(pts, pts, pts.min(0), pts.max(0), 0.1)
    assert "no CUDA device" in str(e.value) or "CUDA" in str(e.value)


def test_bad_arguments_are_reported_not_crashed():
    L = capi.lib()
    h = ctypes.c_void_p()
    pts = np.zeros((4, 3), np.float32)
    rc = L.fgoicp_ctx_create(pts, 0, pts, 4, pts[0], pts[1], 0.1, 0, 0, ctypes.byref(h))
    assert rc == -1 and b"empty" in L.fgoicp_last_error()
    rc = L.fgoicp_ctx_create(pts, 4, pts, 4, pts[0], pts[1], -1.0, 0, 0, ctypes.byref(h))
    assert rc == -1 and b"resolution" in L.fgoicp_last_error()
    assert L.fgoicp_ctx_destroy(None) == 0


def test_preprocess_arguments_and_no_gpu_failure():
    """fgoicp_preprocess: bad arguments reported; without a GPU it fails loudly, never on the CPU."""
    L = capi.lib()
    n = capi.Normalisation()
    pts = np.zeros((4, 3), np.float32)
    assert L.fgoicp_preprocess(pts, 0, pts, 4, 0, 0, ctypes.byref(n)) == -1 and b"empty" in L.fgoicp_last_error()
    assert L.fgoicp_preprocess(pts, 4, pts, 4, 0, 8, ctypes.byref(n)) == -1 and b"flag" in L.fgoicp_last_error()
    assert L.fgoicp_preprocess(pts, 4, pts, 4, 0, 0, None) == -1
    if not torch.cuda.is_available():
        with pytest.raises(capi.FgoicpError) as e:
            capi.preprocess(pts + 1, pts + 2)
        assert "no CUDA device" in str(e.value)
        with pytest.raises(capi.FgoicpError):
            driver.FastGoICP(pts + 1, pts + 2, 0.1, 1e-3, device_preprocess=True)


def test_driver_host_arithmetic_is_bit_identical_to_oracle():
    rng = np.random.default_rng(4)
    for _ in range(300):
        v = (rng.integers(-15, 16, 3) * 0.0625 + rng.choice([0, 0.03125])).astype(np.float32)
        R1, r1 = driver.rotation_matrix(*v)
        R2, r2 = O.rotation(*v)
        assert np.array_equal(R1, R2) and np.float32(r1) == np.float32(r2)
        for span in (0.5, 0.25, 0.125, 0.0625):
            assert bool(driver.overlaps_so3(*v, span)) == O.overlaps_so3(*v, span)
        assert bool(driver.in_so3(*v)) == O.in_so3(*v)


def test_driver_preprocess_is_bit_identical_to_oracle():
    rng = np.random.default_rng(5)
    model = rng.normal(size=(3000, 3)).astype(np.float32) * 40 + 7
    data = rng.normal(size=(700, 3)).astype(np.float32) * 35 - 3
    a, b = driver.preprocess(model, data), O.preprocess(model, data)
    for k in ("model", "data", "offset_pcs", "offset_pct", "bbox_min", "bbox_max"):
        assert np.array_equal(a[k], b[k]), k
    assert np.float32(a["scale"]) == np.float32(b["scale"])
    R, _ = O.rotation(0.3, -0.1, 0.2)
    t = np.array([0.1, -0.2, 0.05], np.float32)
    assert np.array_equal(driver.restore_translation(R, t, a["scale"], a["offset_pcs"], a["offset_pct"]),
                          O.restore_translation(R, t, b["scale"], b["offset_pcs"], b["offset_pct"]))


def test_vectorised_child_enumeration_matches_scalar_rules():
    """The driver's array form of the octant split / overlaps_SO3 / in_SO3 tests equals the scalar rules."""
    F = np.float32
    rng = np.random.default_rng(9)
    for pspan in (1.0, 0.5, 0.25, 0.125):
        pspan = F(pspan)
        span = F(pspan / F(2))
        k = int(round(1 / pspan))
        parents = ((2 * rng.integers(-k, k, size=(40, 3)) + 1) * pspan).astype(F) if pspan < 1 else np.zeros((1, 3), F)
        bits = np.array([[(j >> a) & 1 for a in range(3)] for j in range(8)], F)
        ctr = (parents[:, None, :] - span) + bits[None, :, :] * pspan
        cx, cy, cz = ctr[..., 0].ravel(), ctr[..., 1].ravel(), ctr[..., 2].ravel()
        rr = (cx * cx + cy * cy) + cz * cz
        rq = np.where(rr > F(1.0), rr, np.sqrt(rr)).astype(F)
        a = (np.abs(cx) + np.abs(cy)) + np.abs(cz)
        overl = ((rq - (F(2.0) * span) * a) + (F(3.0) * span) * span) <= F(1.0)
        inside = rq <= F(1.0)
        for i in range(len(cx)):
            assert bool(overl[i]) == O.overlaps_so3(cx[i], cy[i], cz[i], span)
            assert bool(inside[i]) == O.in_so3(cx[i], cy[i], cz[i])

"""Small end-to-end pass over every kernel family for compute-sanitizer (memcheck / racecheck / initcheck)."""
import os, sys, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fast_go_icp_b200 import capi, driver, workloads
w = workloads.synthetic_pair(nt=1500, ns=260, sigma=0.01, seed=3)
pp = driver.preprocess(w["model"], w["data"])
ctx = capi.Context(pp["model"], pp["data"], pp["bbox_min"], pp["bbox_max"], 0.05, flags=capi.BUILD_PACKED | capi.BUILD_TEX)
rot = workloads.rotation_cube_list(140, seed=1)
tc = np.stack([workloads.translation_cube_list(32, level=3, seed=10 + r) for r in range(140)])
tc[::3, 5:9, 3] = -1.0
for phased in (True, False):
    ctx.set_phased(phased)
    ctx.bounds_multi(rot, False, tc)                     # 4480 pairs: phase-ordered kernel when phased
R, _ = driver.rotation_matrix(np.float32(0.1), np.float32(0.2), np.float32(-0.1))
ctx.bounds_batch(R, 0.125, True, tc[0])
ctx.sse(R, np.zeros(3, np.float32)); ctx.nn(R, np.zeros(3, np.float32), True); ctx.nn(R, np.zeros(3, np.float32), False)
ctx.icp(R, np.zeros(3, np.float32), 20, 0.005)
# batches of refinements through both drivers: the persistent loop kernel (mode 2) and the launch chain (mode 1)
R0s = np.stack([driver.rotation_matrix(*rot[k, :3])[0] for k in range(6)]); t0s = np.zeros((6, 3), np.float32)
for mode in (2, 1, 0):
    ctx.set_icp_mode(mode); ctx.icp_batch(R0s, t0s, 12, 0.005)
ctx.bnb_r3_batch(rot[:5], True, 1e10, 260 * 1e-4)
ctx.set_bnb_mode(2); ctx.bnb_r3_batch(rot[:40], False, 3.0, 260 * 1e-4); ctx.set_bnb_mode(0)
ctx.set_trim(0.2)
ctx.bounds_multi(rot[:10], False, tc[:10]); ctx.sse(R, np.zeros(3, np.float32)); ctx.icp(R, np.zeros(3, np.float32), 10, 0.005)
ctx.bnb_r3_batch(rot[:4], True, 1e10, 208 * 1e-4)
ctx.set_trim(0.0)
ctx.close()
g = driver.FastGoICP(w["model"], w["data"], 0.05, 1e-4)
g.run(); g.close()
print("sanitize pass done")

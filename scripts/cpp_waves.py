"""icp::FastGoICP (build/fgoicp_harness) on W5 with FGOICP_DEVICES=0..N-1 and FGOICP_WAVE_LOG=1: per-wave wall time of the
in-process multi-GPU driver against the per-device call and device times.  usage: python scripts/cpp_waves.py <n_devices>"""
import os, subprocess, sys, tempfile
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fast_go_icp_b200 import workloads
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2
w = workloads.synthetic_pair(nt=100_000, ns=10_000, seed=1234)
with tempfile.TemporaryDirectory() as d:
    np.ascontiguousarray(w["model"], np.float32).tofile(os.path.join(d, "m.f32"))
    np.ascontiguousarray(w["data"], np.float32).tofile(os.path.join(d, "d.f32"))
    env = dict(os.environ, FGOICP_DEVICES=",".join(map(str, range(n))), FGOICP_WAVE_LOG="1")
    env.update({k: v for k, v in (a.split("=", 1) for a in sys.argv[2:])})
    r = subprocess.run([os.path.join(ROOT, "build", "fgoicp_harness"), os.path.join(d, "m.f32"), os.path.join(d, "d.f32"), "0.005", "1e-4"],
                       capture_output=True, text=True, env=env)
    print(r.stderr[-6000:])
    res = [l for l in r.stdout.splitlines() if l.startswith("RESULT")][-1].split()[1:]
    v = [float.fromhex(x) for x in res]
    print("run ms %.2f ctor ms %.1f sse %.7g" % (v[15], v[14], v[12]))

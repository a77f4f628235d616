"""Fused (system-resident) vs per-kernel stepping: where do the bits differ, and how fast."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
from triflow_b200 import schemes as S, workloads as W, _lib
from triflow_b200.ensemble import Ensemble
from triflow_b200.model import Model

m = Model(**W.model_args("advdiff"), compiler="cuda")
for sname, kw, N, steps in [("Theta", dict(theta=1), 200, 1), ("ROS2", {}, 200, 1), ("ROS3PRw", dict(time_stepping=False), 200, 1),
                            ("ROS3PRw", dict(time_stepping=False), 4096, 3)]:
    batch = 5
    c = W.ensemble(N, np.arange(0, 32768, 900)[:batch])
    rng = np.random.default_rng(5)
    U0 = np.cos(2 * np.pi * 5 * c["x"]) + 0.3 * rng.standard_normal((batch, N))
    out = []
    for fused in (True, False):
        ens = Ensemble(m, getattr(S, sname)(m, **kw), c["x"], dict(U=U0), c["pars"], hook=S.Dirichlet(U=(1.0, 0.0)), batch=batch)
        ens.set_fusion(fused)
        e = ens.step(c["dt"], steps, want_err=True)
        out.append((ens.download(), e))
    d = np.abs(out[0][0] - out[1][0])
    print(sname, N, "maxdiff", d.max(), "ndiff", (d > 0).sum(), "of", d.size, "err", out[0][1][:3], out[1][1][:3])
    for r in range(min(batch, 2)):
        idx = np.nonzero(d[r])[0]
        print("   member", r, "first/last differing nodes", idx[:6], idx[-6:], "vals", out[0][0][r, idx[:3]], out[1][0][r, idx[:3]])

# speed
members = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
c = W.ensemble(4096, np.arange(members))
sch = S.ROS3PRw(m, time_stepping=False)
import ctypes
for fused in (True, False):
    ens = Ensemble(m, sch, c["x"], c["fields"], c["pars"], hook=S.Dirichlet(U=(1.0, 0.0)), batch=members)
    ens.set_fusion(fused)
    lib, ctx = _lib.lib(), m._cuda.ctx
    ens.step(c["dt"], 3); ens.sync()
    _lib.check(lib.tf_ctx_timer_start(ctx))
    ens.step(c["dt"], 10)
    ms = ctypes.c_float(); _lib.check(lib.tf_ctx_timer_stop(ctx, ctypes.byref(ms)))
    print("fused" if fused else "kernels", "ms/step", ms.value / 10, "node-steps/s %.3e" % (4096 * members * 10 / (ms.value * 1e-3)))
    ens.state.close()

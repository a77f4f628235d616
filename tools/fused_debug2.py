import sys, os
import numpy as np
sys.path.insert(0, ".")
from triflow_b200 import schemes as S, workloads as W, _lib
from triflow_b200.ensemble import Ensemble
from triflow_b200.model import Model
m = Model(**W.model_args("advdiff"), compiler="cuda")
for N in (200, 256, 300, 512, 4096):
    for steps in (1, 2, 3):
        batch = 3
        c = W.ensemble(N, np.arange(0, 32768, 900)[:batch])
        rng = np.random.default_rng(5)
        U0 = np.cos(2 * np.pi * 5 * c["x"]) + 0.3 * rng.standard_normal((batch, N))
        out = []
        for fused in (True, False):
            ens = Ensemble(m, S.ROS3PRw(m, time_stepping=False), c["x"], dict(U=U0), c["pars"], hook=S.Dirichlet(U=(1.0, 0.0)), batch=batch)
            ens.set_fusion(fused)
            for _ in range(steps):
                ens.step(c["dt"], 1)
            out.append(ens.download())
        d = np.abs(out[0] - out[1])
        idx = np.nonzero(d[0])[0]
        print(N, steps, "maxdiff", d.max(), "ndiff", (d > 0).sum(), "first", idx[:5])

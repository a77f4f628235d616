#!/usr/bin/env python
"""Run a few steps of one workload (profiling target)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from triflow_b200 import schemes as S, workloads as W  # noqa
from triflow_b200.ensemble import Ensemble  # noqa
from triflow_b200.model import Model  # noqa

what, N, steps = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
fx = dict(time_stepping=False)
if what == "ks":
    c = W.kuramoto(N)
    m = Model(**W.model_args("ks"), compiler="cuda")
    e = Ensemble(m, S.ROS3PRw(m, **fx), c["x"], c["fields"], c["pars"])
elif what == "ens":
    c = W.ensemble(4096, np.arange(N))
    m = Model(**W.model_args("advdiff"), compiler="cuda")
    e = Ensemble(m, S.ROS3PRw(m, **fx), c["x"], c["fields"], c["pars"],
                 hook=S.Dirichlet(U=(1.0, 0.0)), batch=N)
elif what == "film":
    c = W.film(N)
    m = Model(**W.model_args("film"), compiler="cuda")
    e = Ensemble(m, S.Theta(m), c["x"], c["fields"], c["pars"])
e.step(c["dt"], steps)
e.sync()
print("ok", e.download().sum())

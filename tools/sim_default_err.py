#!/usr/bin/env python
"""Measured deviation of the Simulation-default mode (ROS3PRw double wrapped by the Richardson
controller, reference simulation.py:190-197) from the reference's golden trajectory."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from helpers import rel_traj_err, traj  # noqa: E402
from triflow_b200 import schemes as S, workloads as W  # noqa: E402
from triflow_b200.model import Model  # noqa: E402
from triflow_b200.simulation import Simulation  # noqa: E402

g = traj()
c = W.readme(200)
m = Model(**W.model_args("advdiff"), compiler="cuda")
for hook in (S.Dirichlet(U=(1, 0)), W.readme_hook):
    sim = Simulation(m, dict(x=c["x"], **c["fields"]), c["pars"], dt=c["dt"], tmax=c["tmax"],
                     hook=hook, scheme=S.ROS3PRw)
    snaps = np.array([f.uflat.copy() for _, f in sim])
    ref = g["readme_simdefault_ROS3PRw"]
    per = [rel_traj_err(snaps[i], ref[i]) for i in range(len(ref))]
    print(type(hook).__name__, "rel err per output:", ["%.2e" % e for e in per],
          "sum(U) diff %.3e" % abs(snaps[-1].sum() - 16.777160348256707))

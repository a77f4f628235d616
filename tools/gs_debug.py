#!/usr/bin/env python
"""Where do the grid-resident step, the per-kernel pipeline and the CPU oracle differ?

    python tools/gs_debug.py <case> <N> <steps>
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import gs_check as G  # noqa: E402
from oracle import schemes as O  # noqa: E402
from oracle.numpy_compiler import numpy_compiler  # noqa: E402
from triflow_b200 import workloads as W  # noqa: E402
from triflow_b200.model import Model  # noqa: E402

case, N, steps = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
ug, sg, _ = G.run(case, N, steps, 1)
up, sp, _ = G.run(case, N, steps, 0)
m, sch, c, hook = G.make(case, N)
name = {"ks": "ks", "heat": "heat", "advdiff": "advdiff"}.get(case, None)
uo = None
if name and N <= 20000:
    om = Model(**W.model_args(name), compiler=numpy_compiler)
    f = om.fields_template(x=c["x"], **c["fields"])
    osch = O.ROS3PRw(om, time_stepping=False)
    t = 0.0
    for _ in range(steps):
        t, f = osch(t, f, c["dt"], c["pars"])
    uo = f.uflat
scale = np.max(np.abs(up - up.mean()))
print("status grid %d pipe %d" % (sg, sp))
for lbl, a, b in (("grid-pipe", ug, up), ("grid-oracle", ug, uo), ("pipe-oracle", up, uo)):
    if b is None:
        continue
    d = np.abs(a - b) / scale
    idx = np.argsort(d)[-8:][::-1]
    print("%-12s max %.3e  worst nodes %s" % (lbl, d.max(), [(int(i), float("%.2e" % d[i])) for i in idx]))

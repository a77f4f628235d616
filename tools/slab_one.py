#!/usr/bin/env python
"""A few steps of the several-GPU step kernel (tf_k_gridstep_mr) with ONE slab on one GPU, for
ncu (a multi-rank run cannot be profiled: ncu replays kernels, the ranks wait for each other).

    python tools/slab_one.py [N] [steps]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from triflow_b200 import distributed as D, schemes as S, workloads as W  # noqa: E402
from triflow_b200.model import Model  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 6
c = W.kuramoto(N)
m = Model(**W.model_args("ks"), compiler="cuda")
g = D.SlabGrid(m, S.ROS3PRw(m, time_stepping=False), c["x"], c["fields"], c["pars"], devices=[0])
for _ in range(steps):
    g.step(c["dt"], 1)
u = g.gather()
print("slab of %d nodes, %d tiles, %d steps, sum(U) = %.12g" % (N, g.states[0].tiles_total, steps, u.sum()))
g.close()

#!/usr/bin/env python
"""Where does the time of the device-resident output path go? (host-side segments)"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from triflow_b200 import _lib, schemes as S, simulation as SIM, workloads as W  # noqa: E402
from triflow_b200.model import Model  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 18
c = W.film(N)
m = Model(**W.model_args("film"), compiler="cuda")
steps = 40
kw = dict(dt=c["dt"], tmax=steps * c["dt"], scheme=S.Theta, time_stepping=False)
acc = {}


def timed(name, fn):
    def w(*a, **k):
        t0 = time.perf_counter()
        r = fn(*a, **k)
        acc[name] = acc.get(name, 0.0) + time.perf_counter() - t0
        return r
    return w


SIM.OutputRing.push = timed("ring.push", SIM.OutputRing.push)
SIM.OutputRing.pop = timed("ring.pop(consumer)", SIM.OutputRing.pop)
SIM.Simulation._ring_push = timed("_ring_push", SIM.Simulation._ring_push)
SIM.Simulation._compute_one_step = timed("_compute_one_step", SIM.Simulation._compute_one_step)
lib = _lib.lib()
orig_push = lib.tf_ring_push


def run(ring):
    acc.clear()
    sim = SIM.Simulation(m, dict(x=c["x"], **c["fields"]), c["pars"], ring=ring, lazy=True, **kw)
    sim.stream.sink(lambda fr: None)
    t0 = time.perf_counter()
    for _ in sim:
        pass
    _lib.check(lib.tf_ctx_sync(m._cuda.ctx))
    return time.perf_counter() - t0


run(0)
for ring in (0, 4, 0, 4, 16):
    t = run(ring)
    print("ring=%d total %.1f ms (%.2f ms/step)" % (ring, t * 1e3, t * 1e3 / steps),
          {k: "%.1f ms" % (v * 1e3) for k, v in acc.items()}, flush=True)

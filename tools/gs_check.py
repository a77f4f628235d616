#!/usr/bin/env python
"""Grid-resident step (tf_k_gridstep) against the per-kernel pipeline: agreement and time.

    python tools/gs_check.py [quick|full|time]

Prints one line per case: max relative difference after a few steps, status, ms/step of both
paths.  Every case runs in this process; the kernel's waits are bounded (status bit 4).
"""
import ctypes
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from triflow_b200 import _lib, schemes as S, workloads as W  # noqa: E402
from triflow_b200.ensemble import Ensemble  # noqa: E402
from triflow_b200.model import Model  # noqa: E402

_models = {}


def model(name):
    if name not in _models:
        _models[name] = Model(**W.model_args(name), compiler="cuda")
    return _models[name]


def make(case, N):
    fx = dict(time_stepping=False)
    if case == "ks":
        c = W.kuramoto(N)
        m = model("ks")
        return m, S.ROS3PRw(m, **fx), c, S.null_hook
    if case == "ks_ros2":
        c = W.kuramoto(N)
        m = model("ks")
        return m, S.ROS2(m), c, S.null_hook
    if case == "ks_theta":
        c = W.kuramoto(N)
        m = model("ks")
        return m, S.Theta(m, theta=0.5), c, S.null_hook
    if case.startswith("burgers"):
        acc = int(case[-1])
        c = W.burgers(N, acc)
        m = model("burgers_up%d" % acc)
        return m, S.ROS2(m), c, S.null_hook
    if case == "advdiff":          # non-periodic, Dirichlet hook
        c = W.readme(N)
        c["dt"] = 0.01
        m = model("advdiff")
        return m, S.ROS3PRw(m, **fx), c, S.Dirichlet(U=(1.0, 0.0))
    if case == "advdiff_node":     # periodic, per-node parameter, slowly decaying propagator
        rng = np.random.default_rng(3)
        x = np.linspace(0, 10, N)
        U = np.cos(2 * np.pi * x / 10) + 0.1 * rng.standard_normal(N)
        c = dict(x=x, fields=dict(U=U), pars=dict(k=1e-2 * (1 + rng.random(N)), c=.3, periodic=True),
                 dt=0.01)
        m = model("advdiff")
        return m, S.ROS3PRw(m, **fx), c, S.null_hook
    if case == "heat":             # periodic, large a: the border fill reaches every tile
        x = np.linspace(0, 10, N)
        c = dict(x=x, fields=dict(T=np.cos(2 * np.pi * x / 10)), pars=dict(k=1.0, periodic=True), dt=0.5)
        m = model("heat")
        return m, S.ROS3PRw(m, **fx), c, S.null_hook
    if case == "ks_edge":          # non-periodic pentadiagonal
        c = W.kuramoto(N)
        c["pars"] = dict(periodic=False)
        m = model("ks")
        return m, S.ROS3PRw(m, **fx), c, S.null_hook
    raise SystemExit(case)


def run(case, N, steps, fuse, timing=0):
    m, sch, c, hook = make(case, N)
    e = Ensemble(m, sch, c["x"], c["fields"], c["pars"], hook=hook, batch=1)
    e.set_fusion("grid" if fuse else 0)
    lib, ctx = _lib.lib(), m._cuda.ctx
    e.step(c["dt"], steps)
    e.sync()
    u = e.download()[0].copy()
    ms = None
    if timing:
        e.step(c["dt"], 3)
        e.sync()
        _lib.check(lib.tf_ctx_timer_start(ctx))
        e.step(c["dt"], timing)
        t = ctypes.c_float()
        _lib.check(lib.tf_ctx_timer_stop(ctx, ctypes.byref(t)))
        ms = t.value / timing
    st = e.state.status()
    e.state.close()
    return u, int(st[0]), ms


def check(case, N, steps, timing=0):
    t0 = time.time()
    try:
        ug, sg, tg = run(case, N, steps, 1, timing)
    except Exception as ex:  # noqa: BLE001
        print("%-10s N=%-8d GRIDSTEP FAILED: %s" % (case, N, str(ex)[:150]), flush=True)
        return
    up, sp, tp = run(case, N, steps, 0, timing)
    scale = np.max(np.abs(up - up.mean())) or 1.0
    d = np.max(np.abs(ug - up)) / scale
    where = int(np.argmax(np.abs(ug - up)))
    print("%-10s N=%-8d steps=%-3d rel diff %.2e (at %d) status %d/%d finite %s  ms/step grid %s pipe %s  [%.1fs]"
          % (case, N, steps, d, where, sg, sp, bool(np.isfinite(ug).all()),
             "%.4f" % tg if tg else "-", "%.4f" % tp if tp else "-", time.time() - t0), flush=True)


def main():
    mode = sys.argv[1] if len(sys.argv) > 1 else "quick"
    if mode in ("quick", "full"):
        check("ks", 4096, 3)
        check("ks", 1 << 14, 3)
        check("ks", 50000, 3)
        check("ks", 1 << 17, 3)
        check("burgers1", 1 << 14, 3)
        check("advdiff", 20000, 3)
        check("ks_edge", 30000, 3)
        check("advdiff_node", 3000, 3)
        check("heat", 5000, 3)
        check("heat", 1 << 16, 3)
    if mode == "full":
        check("ks", 1000, 5)
        check("ks", 2049, 5)
        check("ks", 123457, 5)
        check("ks_ros2", 70001, 5)
        check("ks_theta", 70001, 5)
        check("burgers2", 1 << 16, 5)
        check("burgers3", 99999, 5)
        check("advdiff", 5000, 5)
        check("ks", 1 << 20, 10)
    if mode == "sweep":
        for case in ("ks", "burgers1"):
            for e in (10, 12, 14, 16, 17, 18, 19, 20):
                check(case, 1 << e, 3, timing=20)
    if mode in ("quick", "full", "time"):
        check("ks", 1 << 20, 3, timing=20)
        check("burgers1", 1 << 17, 3, timing=20)
        check("ks", 1 << 18, 3, timing=20)


if __name__ == "__main__":
    main()

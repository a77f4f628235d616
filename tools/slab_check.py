#!/usr/bin/env python
"""One grid over several GPUs (triflow_b200.distributed.SlabGrid) against the same grid stepped on
one GPU: agreement and time.

    python tools/slab_check.py local 0,1 [N ...]        one process drives the listed GPUs
    python tools/slab_check.py local 0,0 [N ...]        two slabs on ONE GPU (protocol check)
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tools/slab_check.py dist [N ...]                one process per GPU, IPC-mapped record areas

Prints one line per case (rank 0): max relative difference after a few steps against the
single-GPU result, ms/step of both.
"""
import ctypes
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from triflow_b200 import _lib, distributed as D, schemes as S, workloads as W  # noqa: E402
from triflow_b200.ensemble import Ensemble  # noqa: E402
from triflow_b200.model import Model  # noqa: E402


def make(case, N):
    fx = dict(time_stepping=False)
    if case == "ks":
        m = Model(**W.model_args("ks"), compiler="cuda")
        return m, S.ROS3PRw(m, **fx), W.kuramoto(N)
    if case == "ks_edge":
        m = Model(**W.model_args("ks"), compiler="cuda")
        c = W.kuramoto(N)
        c["pars"] = dict(periodic=False)
        return m, S.ROS3PRw(m, **fx), c
    if case == "heat":             # periodic, large a: the border fill reaches every tile
        m = Model(**W.model_args("heat"), compiler="cuda")
        x = np.linspace(0, 10, N)
        c = dict(x=x, fields=dict(T=np.cos(2 * np.pi * x / 10)), pars=dict(k=1.0, periodic=True), dt=0.5)
        return m, S.ROS3PRw(m, **fx), c
    if case == "burgers":
        m = Model(**W.model_args("burgers_up1"), compiler="cuda")
        return m, S.ROS2(m), W.burgers(N, 1)
    if case == "advdiff":          # non-periodic, Dirichlet hook at both ends
        m = Model(**W.model_args("advdiff"), compiler="cuda")
        c = W.readme(N)
        c["dt"] = 0.01
        c["hook"] = S.Dirichlet(U=(1.0, 0.0))
        return m, S.ROS3PRw(m, **fx), c
    raise SystemExit(case)


def single(m, sch, c, steps, timing):
    e = Ensemble(m, sch, c["x"], c["fields"], c["pars"], hook=c.get("hook", S.null_hook), batch=1)
    e.step(c["dt"], steps)
    e.sync()
    u = e.download()[0].copy()
    ms = None
    if timing:
        lib, ctx = _lib.lib(), m._cuda.ctx
        e.step(c["dt"], 3)
        e.sync()
        _lib.check(lib.tf_ctx_timer_start(ctx))
        e.step(c["dt"], timing)
        t = ctypes.c_float()
        _lib.check(lib.tf_ctx_timer_stop(ctx, ctypes.byref(t)))
        ms = t.value / timing
    e.state.close()
    return u, ms


def slab(m, sch, c, steps, timing, devices):
    g = D.SlabGrid(m, sch, c["x"], c["fields"], c["pars"], devices=devices, hook=c.get("hook"))
    try:
        if os.environ.get("SLAB_DEBUG"):
            g.step(c["dt"], 1)
            g.sync()
        g.step(c["dt"], steps - (1 if os.environ.get("SLAB_DEBUG") else 0))
        u = g.gather()
    except RuntimeError:
        for s in g.states:          # where every tile was waiting when the launch gave up
            stuck = (ctypes.c_int * 1024)()
            _lib.lib().tf_model_read_symbol(s.variant.handle, b"tf_gs_stuck", stuck, 4096)
            when = (ctypes.c_ulonglong * 2)()
            _lib.lib().tf_model_read_symbol(s.variant.handle, b"tf_gs_when", when, 16)
            print("rank %d status %#x start %d first time-out +%.3f ms; first wait site per tile: %s" % (
                s.rank, s.status(), when[0], (when[1] - when[0]) * 1e-6, [hex(v) for v in stuck[:s.tiles_local]]),
                flush=True)
        raise
    ms = None
    if timing:
        g.step(c["dt"], 3)
        g.sync()
        D.barrier()
        t0 = time.perf_counter()
        g.step(c["dt"], timing)
        g.sync()
        ms = D.max_over_ranks((time.perf_counter() - t0) * 1e3 / timing)
    info = (g.states[0].tiles_local, g.states[0].tiles_total, g.nranks)
    g.close()
    return u, ms, info


def check(case, N, steps, devices, timing=0):
    t0 = time.time()
    m, sch, c = make(case, N)
    rank = 0 if devices is not None else D.world()[0]
    try:
        ug, tg, info = slab(m, sch, c, steps, timing, devices)
    except Exception as ex:  # noqa: BLE001
        print("[rank %d] %-8s N=%-8d SLAB FAILED: %s" % (rank, case, N, str(ex)[:200]), flush=True)
        return False
    if rank != 0:
        return True
    up, tp = single(m, sch, c, steps, timing)
    scale = np.max(np.abs(up - up.mean())) or 1.0
    d = np.max(np.abs(ug - up)) / scale
    ok = bool(d < 1e-10 and np.isfinite(ug).all())
    print("%-8s N=%-8d steps=%-3d ranks %d tiles %d/rank (%d live)  rel diff %.2e  ms/step slab %s single %s  %s [%.1fs]"
          % (case, N, steps, info[2], info[0], info[1], d, "%.4f" % tg if tg else "-",
             "%.4f" % tp if tp else "-", "ok" if ok else "MISMATCH", time.time() - t0), flush=True)
    return ok


def main():
    mode = sys.argv[1]
    if mode == "local":
        devices = [int(d) for d in sys.argv[2].split(",")]
        sizes = sys.argv[3:]
    else:
        devices = None
        D.init()
        sizes = sys.argv[2:]
    n = len(devices) if devices else D.world()[1]
    ok = True
    if not sizes:
        for case, N in (("ks", 20000), ("ks", 50001), ("ks_edge", 30000), ("heat", 20000), ("burgers", 40000),
                        ("advdiff", 30000)):
            ok &= check(case, N, 3, devices)
        if not (devices and len(set(devices)) < len(devices)):
            ok &= check("ks", n << 20, 3, devices, timing=20)
    for a in sizes:                       # N or case:N
        case, N = a.split(":") if ":" in a else ("ks", a)
        ok &= check(case, int(N), 3, devices, timing=20 if int(N) >= 1 << 18 else 0)
    if devices is None:
        D.finalize()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()

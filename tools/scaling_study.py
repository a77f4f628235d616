#!/usr/bin/env python
"""Per-kernel-family device times vs grid size (tuning aid, not a test)."""
import ctypes
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from triflow_b200 import _lib, schemes as S, workloads as W  # noqa
from triflow_b200.compiler import make_cuda_compiler  # noqa
from triflow_b200.ensemble import Ensemble  # noqa
from triflow_b200.model import Model  # noqa


def run(name, cfg, mk, steps=10, batch=None, chunk=None, fast_div=True):
    m = Model(**W.model_args(name), compiler=make_cuda_compiler(chunk_nodes=chunk, fast_div=fast_div))
    sch = mk(m)
    ens = Ensemble(m, sch, cfg["x"], cfg["fields"], cfg["pars"],
                   hook=S.Dirichlet(**{k: v for k, v in cfg.get("dirichlet", {}).items()})
                   if cfg.get("dirichlet") else S.null_hook, batch=batch)
    lib, ctx = _lib.lib(), m._cuda.ctx
    ens.step(cfg["dt"], 3)
    ens.sync()
    ms = ctypes.c_float()
    lib.tf_ctx_timer_start(ctx)
    ens.step(cfg["dt"], steps)
    lib.tf_ctx_timer_stop(ctx, ctypes.byref(ms))
    lib.tf_ctx_profile(ctx, 1)
    ens.step(cfg["dt"], 3)
    fam = {}
    for i, f in enumerate(_lib.FAMILIES):
        a, n = ctypes.c_float(), ctypes.c_longlong()
        lib.tf_ctx_profile_read(ctx, i, ctypes.byref(a), ctypes.byref(n))
        if n.value:
            fam[f] = round(a.value / n.value * 1e3, 1)
    lib.tf_ctx_profile(ctx, 0)
    units = cfg["x"].size * ens.batch
    print("%-10s N=%-8d batch=%-6d M=%s NW=%s  %.1f us/step  %.3e node-steps/s  per-launch us: %s" % (
        name, cfg["x"].size, ens.batch, chunk, os.environ.get("TF_NW", "-"),
        ms.value / steps * 1e3, units * steps / (ms.value * 1e-3), fam), flush=True)
    ens.state.close()


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "ks"
    fx = dict(time_stepping=False)
    if what == "ks":
        for N in [2 ** 12, 2 ** 14, 2 ** 16, 2 ** 18, 2 ** 20, 2 ** 22, 2 ** 24]:
            run("ks", W.kuramoto(N), lambda m: S.ROS3PRw(m, **fx))
    elif what == "ksm":
        for chunk in (4, 8, 16):
            run("ks", W.kuramoto(2 ** 20), lambda m: S.ROS3PRw(m, **fx), chunk=chunk)
    elif what == "ens":
        for nb in (1024, 8192, 32768):
            run("advdiff", W.ensemble(4096, np.arange(nb)), lambda m: S.ROS3PRw(m, **fx), batch=nb)
    elif what == "ensm":
        for chunk in (4, 8, 16):
            run("advdiff", W.ensemble(4096, np.arange(8192)), lambda m: S.ROS3PRw(m, **fx),
                batch=8192, chunk=chunk)
    elif what == "ksens":
        # 2^20 nodes as independent KS systems (no look-back) vs one long grid
        for N, nb in ((4096, 256), (3072, 342), (2048, 512), (65536, 16)):
            c = W.kuramoto(N)
            run("ks", c, lambda m: S.ROS3PRw(m, **fx), batch=nb)
    elif what == "film":
        run("film", W.film(2 ** 18), lambda m: S.Theta(m, theta=1))

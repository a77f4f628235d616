#!/usr/bin/env python
"""Per-tile phase timeline of the grid-resident step (cubin built with -DTF_GS_TRACE).

    TF_CFLAGS=-DTF_GS_TRACE python tools/gs_trace.py [ks|burgers] [N]
"""
import ctypes
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from triflow_b200 import _lib, schemes as S, workloads as W  # noqa: E402
from triflow_b200.ensemble import Ensemble  # noqa: E402
from triflow_b200.model import Model  # noqa: E402

what = sys.argv[1] if len(sys.argv) > 1 else "ks"
N = int(sys.argv[2]) if len(sys.argv) > 2 else (1 << 20 if what == "ks" else 1 << 17)
if what == "ks":
    c, name, mk, s = W.kuramoto(N), "ks", lambda m: S.ROS3PRw(m, time_stepping=False), 3
else:
    c, name, mk, s = W.burgers(N, 1), "burgers_up1", lambda m: S.ROS2(m), 2
m = Model(**W.model_args(name), compiler="cuda")
e = Ensemble(m, mk(m), c["x"], c["fields"], c["pars"], batch=1)
e.set_fusion("grid")
e.step(c["dt"], 6)
e.sync()
buf = np.zeros(512 * 32, dtype=np.uint64)
_lib.check(_lib.lib().tf_model_read_symbol(e.state.variant.handle, b"tf_gs_trace",
                                           buf.ctypes.data_as(ctypes.c_void_p), buf.nbytes))
tr = buf.reshape(512, 32).astype(np.int64)
tiles = int((tr[:, 0] > 0).sum())
tr = tr[:tiles]
t0 = tr[:, 0].min()
names = ["start", "factor pass1", "factor scan", "factor pass2+L"]
for i in range(s):
    names += ["st%d state+F+fwd1" % i, "st%d fwd scan" % i, "st%d fwd2+border" % i,
              "st%d bwd1" % i, "st%d bwd scan" % i, "st%d bwd2+out" % i]
print("%s N=%d tiles=%d; step span %.1f us (first start -> last end)" % (
    what, N, tiles, (tr[:, 3 + 6 * s].max() - t0) / 1e3))
print("%-22s %8s %8s %8s   %s" % ("phase", "med us", "max us", "min us", "end (median, us from start)"))
for k in range(1, 4 + 6 * s):
    d = (tr[:, k] - tr[:, k - 1]) / 1e3
    print("%-22s %8.2f %8.2f %8.2f   %8.2f" % (names[k], np.median(d), d.max(), d.min(),
                                               np.median(tr[:, k] - t0) / 1e3))
print("start skew: %.2f us; slowest tiles by end time: %s" % (
    (tr[:, 0].max() - t0) / 1e3, np.argsort(tr[:, 3 + 6 * s])[-5:].tolist()))
if tr[:, 22].max() > 0:
    late = (tr[:, 22] - tr[:, 1]) / 1e3
    print("factor pass 1, latest thread of a tile behind thread 0: med %.2f max %.2f us (tile %d)" % (
        np.median(late), late.max(), int(late.argmax())))
if len(sys.argv) > 3:
    np.save(sys.argv[3], tr - t0)          # per-tile stamps (ns from the first start) for offline analysis

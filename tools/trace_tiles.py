"""Per-tile phase timeline of the chained sweep kernels (cubin built with -DTF_TRACE).
Run under gpurun:  TF_CFLAGS=-DTF_TRACE python tools/trace_tiles.py [workload] [nodes]"""
import ctypes as C
import os
import sys
import numpy as np
sys.path.insert(0, ".")
assert "TF_TRACE" in os.environ.get("TF_CFLAGS", ""), "set TF_CFLAGS=-DTF_TRACE"
import bench
from triflow_b200 import _lib, workloads as W
from triflow_b200.ensemble import Ensemble
from triflow_b200.model import Model

wl = sys.argv[1] if len(sys.argv) > 1 else "ks"
nodes = int(sys.argv[2]) if len(sys.argv) > 2 else None
mname, mk, x, fields, pars, hook, batch, N, dt = bench.build_problem(wl, None, nodes)
model = Model(**W.model_args(mname), compiler="cuda")
sch = mk(model)
ens = Ensemble(model, sch, x, fields, pars, hook=hook, batch=batch)
os.environ["TF_NO_GRAPH"] = "1"
ens.step(dt, 3)
ens.sync()
ens.step(dt, 2)
ens.sync()
buf = np.zeros(16 * 4096 * 8, dtype=np.uint64)
h = ens.state.variant.handle
_lib.check(_lib.lib().tf_model_read_symbol(h, b"tf_trace", buf.ctypes.data_as(C.c_void_p), buf.nbytes))
tr = buf.reshape(16, 4096, 8).astype(np.int64)
buf2 = np.zeros(16 * 4096 * 8, dtype=np.uint64)
_lib.check(_lib.lib().tf_model_read_symbol(h, b"tf_trace2", buf2.ctypes.data_as(C.c_void_p), buf2.nbytes))
tr2 = buf2.reshape(16, 4096, 8).astype(np.int64)
names = ["entry", "loaded", "tma", "pass1", "lb0", "lb1", "scanned", "end"]
for slot in range(16):
    t = tr[slot]
    used = np.nonzero(t[:, 0])[0]
    if used.size == 0:
        continue
    t = t[used]
    t0 = t[:, 0].min()
    rel = (t - t0) / 1e3
    print("slot %2d: %4d tiles; entry spread %.1f us; kernel span %.1f us" % (
        slot, used.size, rel[:, 0].max(), rel[:, 7].max()))
    d = np.diff(t, axis=1) / 1e3
    print("   phase durations (us) median/max: " + "  ".join(
        "%s->%s %.1f/%.1f" % (names[i], names[i + 1], np.median(d[:, i]), d[:, i].max()) for i in range(7)))
    rd = tr[slot][used][:, 2]
    rounds, depth = rd // 1000, rd % 1000
    print("   look-back rounds median/max %d/%d   tiles combined median/max %d/%d" % (
        np.median(rounds), rounds.max(), np.median(depth), depth.max()))
    t[:, 2] = t[:, 1]
    rel = (t - t0) / 1e3
    lb0, lb1 = rel[:, 4], rel[:, 5]
    for k in (1, 4, 8, 32):
        mx = np.array([lb0[max(0, i - k):i + 1].max() for i in range(len(lb0))])
        print("   lb1 - max(lb0 of self and %2d predecessors): median %.1f  p90 %.1f  max %.1f" % (
            k, np.median(lb1 - mx), np.percentile(lb1 - mx, 90), (lb1 - mx).max()))
    print("   lb0 percentiles 10/50/90/100: %.1f %.1f %.1f %.1f" % tuple(np.percentile(lb0, [10, 50, 90, 100])))
    if slot in (0, 2):
        slow = [i for i in range(150, len(lb0)) if lb1[i] - lb0[i] > 10][:8]
        for i in slow:
            r2 = tr2[slot][used[i]]
            print("   SLOW tile %d lb0=%.1f lb1=%.1f rounds: " % (used[i], lb0[i], lb1[i]) + " ".join(
                "t=%.1f kr/hit/abs=%d" % ((r2[q] - t0) / 1e3, r2[4 + q]) for q in range(min(4, rounds[i]))) +
                "  | predecessors (lb0,lb1): " + " ".join("(%.0f,%.0f)" % (lb0[i - q], lb1[i - q]) for q in range(1, 6)))
    if slot in (0, 2):
        print("   tiles 200..239 (lb0,lb1): " + " ".join("(%.0f,%.0f,r%d,d%d)" % (lb0[i], lb1[i], rounds[i], depth[i]) for i in range(200, min(240, len(lb0)))))
    for q in (0, 1, 2, len(used) // 2, len(used) - 1):
        print("   tile %4d: " % used[q] + " ".join("%s=%.1f" % (names[i], rel[q, i]) for i in range(8)))

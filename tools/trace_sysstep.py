"""Phase clocks of the system-resident kernel (cubin built with -DTF_TRACE): cycles of thread 0
of every CTA per phase, summed over the systems the CTA stepped.
Run under gpurun:  TF_CFLAGS=-DTF_TRACE python tools/trace_sysstep.py [members]"""
import ctypes as C
import os
import sys
import numpy as np
sys.path.insert(0, ".")
assert "TF_TRACE" in os.environ.get("TF_CFLAGS", ""), "set TF_CFLAGS=-DTF_TRACE"
from triflow_b200 import _lib, schemes as S, workloads as W
from triflow_b200.ensemble import Ensemble
from triflow_b200.model import Model

members = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
m = Model(**W.model_args("advdiff"), compiler="cuda")
c = W.ensemble(4096, np.arange(members) % 32768)
ens = Ensemble(m, S.ROS3PRw(m, time_stepping=False), c["x"], c["fields"], c["pars"],
               hook=S.Dirichlet(U=(1.0, 0.0)), batch=members)
os.environ["TF_NO_GRAPH"] = "1"
steps = 4
ens.step(c["dt"], steps)
ens.sync()
buf = np.zeros(16 * 4096 * 8, dtype=np.uint64)
_lib.check(_lib.lib().tf_model_read_symbol(ens.state.variant.handle, b"tf_trace",
                                           buf.ctypes.data_as(C.c_void_p), buf.nbytes))
tr = buf[:1024 * 32].reshape(1024, 32).astype(np.float64)[:148]
names = {12: "  hand-over: error estimate out", 13: "  hand-over: next TMA issued", 14: "  hand-over: constants / pointers",
         15: "  hand-over: mbarrier wait (U arrived)", 0: "  hand-over: barrier", 5: "factor: edge-row pre-pass", 6: "factor: pass 1", 7: "factor: scan",
         1: "factor: pass 2", 2: "border block", 8: "stage: state/halo/F/fwd pass 1 (x3)",
         9: "stage: fwd scan (x3)", 10: "stage: fwd pass 2 + border + bwd pass 1 (x3)",
         11: "stage: bwd scan (x3)", 3: "stage: bwd pass 2 (+update) (x3)", 4: "error reduce + end barrier"}
nsys = members * steps / 148.0
tot = tr.sum(axis=1).mean()
print("cycles per system and step: %.0f (%.1f us at 1965 MHz)" % (tot / nsys, tot / nsys / 1965.0))
for k in (12, 13, 14, 15, 0, 5, 6, 7, 1, 2, 8, 9, 10, 11, 3, 4):
    v = tr[:, k].mean() / nsys
    print("  %-48s %7.0f cycles  %5.1f %%" % (names[k], v, 100 * v * nsys / tot))

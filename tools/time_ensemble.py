"""Device-resident timing of the config-5 ensemble step (no checks): quick A/B of kernel variants.
usage: [TF_CFLAGS=...] python tools/time_ensemble.py [members] [steps]"""
import ctypes
import sys
import numpy as np
sys.path.insert(0, ".")
from triflow_b200 import schemes as S, workloads as W, _lib
from triflow_b200.ensemble import Ensemble
from triflow_b200.model import Model

members = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
m = Model(**W.model_args("advdiff"), compiler="cuda")
c = W.ensemble(4096, np.arange(members) % 32768)
ens = Ensemble(m, S.ROS3PRw(m, time_stepping=False), c["x"], c["fields"], c["pars"],
               hook=S.Dirichlet(U=(1.0, 0.0)), batch=members)
lib, ctx = _lib.lib(), m._cuda.ctx
lib.tf_scheme_step(ens.state.h, ens.scheme.handle, float(c["dt"]), 3, None)
ens.sync()
_lib.check(lib.tf_ctx_timer_start(ctx))
lib.tf_scheme_step(ens.state.h, ens.scheme.handle, float(c["dt"]), steps, None)
ms = ctypes.c_float()
_lib.check(lib.tf_ctx_timer_stop(ctx, ctypes.byref(ms)))
print("ms/step %.4f  grid-point*steps/s %.4e" % (ms.value / steps, 4096.0 * members * steps / (ms.value * 1e-3)))

#!/usr/bin/env python
"""Small end-to-end cases for compute-sanitizer (memcheck / racecheck)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from triflow_b200 import schemes as S, workloads as W  # noqa
from triflow_b200.ensemble import Ensemble  # noqa
from triflow_b200.model import Model  # noqa

fx = dict(time_stepping=False)
# periodic, several look-back tiles, ragged N
c = W.kuramoto(9001)
m = Model(**W.model_args("ks"), compiler="cuda")
f = m.fields_template(x=c["x"], **c["fields"])
t, f = S.ROS3PRw(m, **fx).run_fixed(0.0, f, c["dt"], 2, c["pars"])
print("ks", float(f.uflat.sum()))
m.F(f, c["pars"]); m.J(f, c["pars"])
# non-periodic ensemble with Dirichlet hook
ce = W.ensemble(700, np.arange(5))
ma = Model(**W.model_args("advdiff"), compiler="cuda")
e = Ensemble(ma, S.ROS3PRw(ma, **fx), ce["x"], ce["fields"], ce["pars"],
             hook=S.Dirichlet(U=(1.0, 0.0)), batch=5)
e.step(ce["dt"], 2)
print("ens", float(e.download().sum()))
# adaptive controller
cr = W.readme(200)
fr = ma.fields_template(x=cr["x"], **cr["fields"])
t, fr = S.ROS3PRw(ma, tol=1e-1)(0.0, fr, 0.5, cr["pars"], hook=S.Dirichlet(U=(1, 0)))
print("adaptive", float(fr.uflat.sum()))
if "film" in sys.argv:
    cf = W.film(600)
    mf = Model(**W.model_args("film"), compiler="cuda")
    ff = mf.fields_template(x=cf["x"], **cf["fields"])
    t, ff = S.Theta(mf)(0.0, ff, cf["dt"], cf["pars"])
    print("film", float(ff.uflat.sum()))

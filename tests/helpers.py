"""Shared helpers for the parity tests (oracle side)."""
import json
import os

import numpy as np
from scipy.sparse import csc_matrix

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_expr():
    with open(os.path.join(GOLDEN, "expr.json")) as f:
        return json.load(f)


def load_fj(tag):
    d = np.load(os.path.join(GOLDEN, "fj_%s.npz" % tag))
    fields = {k[6:]: d[k] for k in d.files if k.startswith("field_")}
    pars = {k[4:]: (d[k] if d[k].ndim else float(d[k]))
            for k in d.files if k.startswith("par_")}
    pars["periodic"] = bool(d["periodic"])
    n = d["F"].size
    J = csc_matrix((d["J_data"], d["J_indices"], d["J_indptr"]), shape=(n, n))
    return d["x"], fields, pars, d["F"], J


def fj_tags():
    return sorted(f[3:-4] for f in os.listdir(GOLDEN)
                  if f.startswith("fj_") and f.endswith(".npz"))


def model_name_of(tag):
    from triflow_b200 import workloads as W
    alias = {"helperdx": "helper_dx", "upwindconst": "upwind_const"}
    for name in sorted(W.MODELS, key=len, reverse=True):
        pass
    base = tag
    for suffix in ("_per", "_edge"):
        if base.endswith(suffix):
            base = base[: -len(suffix)]
    for strip in ("_arrpar",):
        base = base.replace(strip, "")
    if base.startswith("ks_tiny"):
        base = "ks"
    return alias.get(base, base)


def traj():
    return np.load(os.path.join(GOLDEN, "traj.npz"))


def traj_ens4096():
    """cfg 5 at its real grid size: 64 members x 100 steps run by the reference itself."""
    return np.load(os.path.join(GOLDEN, "traj_ens4096.npz"))


def csc_triplet(J):
    J = J.tocsc().copy()
    J.sum_duplicates()
    J.sort_indices()
    return J.indptr, J.indices, J.data


def rel_traj_err(U, Uref):
    scale = np.max(np.abs(Uref - Uref.mean()))
    return np.max(np.abs(U - Uref)) / scale

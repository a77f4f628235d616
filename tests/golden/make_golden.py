#!/usr/bin/env python
"""Generate the golden fixtures in this directory FROM THE REFERENCE ITSELF.

Run in the build container only (``/root/reference`` must be mounted):

    python tests/golden/make_golden.py

The reference's ``triflow/core/{compilers,routines,fields,model,schemes,
simulation}.py`` are imported *unmodified, by file path* under the import shims
of SURVEY.md Appendix A (xarray / toolz / pendulum / streamz / path are not
installed), with the two compatibility patches of SURVEY.md §8c:

1. SymPy 1.14 prints ``Heaviside(x, 1/2)``; the reference's one-argument
   override (``compilers.py:204-205``) would raise ``TypeError`` -> the NumPy
   printer is told to print the one-argument form, so the reference's own
   ``np_Heaviside`` (always 1) is what runs.
2. ``BaseFields.uflat`` indexes with a list of slices (``fields.py:154-157``),
   removed in numpy >= 1.23 -> a duck-typed Fields stand-in with the same
   layout (``uflat[i*nvar+e]``) is used instead of the xarray subclass.

Outputs (committed): ``expr.json`` (expression trees, printed source, model
attributes), ``fj_*.npz`` (F and J triplets), ``traj_*.npz`` (trajectories and
controller traces).  Nothing here is imported by the product.
"""

import importlib.util
import inspect
import json
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("TRIFLOW_REFERENCE", "/root/reference")
sys.path.insert(0, ROOT)

from triflow_b200 import workloads as W  # noqa: E402  (inputs only)


# ----------------------------------------------------------------- shims
class _RefArray(np.ndarray):
    @property
    def values(self):
        return self.view(np.ndarray)


class RefFields:
    """Duck-typed stand-in for the reference's xarray-based BaseFields."""

    def __init__(self, deps, helps, **inputs):
        self.dependent_variables = list(deps)
        self.helper_functions = list(helps)
        self._d = {k: np.array(inputs[k], dtype=float).view(_RefArray)
                   for k in ["x", *deps, *helps]}

    def __getitem__(self, k):
        return self._d[k]

    def __setitem__(self, k, v):
        self._d[k][...] = v

    @property
    def size(self):
        return self._d["x"].size

    @property
    def uflat(self):
        return np.vstack([self._d[k].view(np.ndarray)
                          for k in self.dependent_variables]).flatten("F")

    def fill(self, uflat):
        r = np.asarray(uflat).reshape((self.size, -1))
        for e, k in enumerate(self.dependent_variables):
            self._d[k][...] = r[:, e]

    def copy(self, deep=True):
        return RefFields(self.dependent_variables, self.helper_functions, **self._d)


def load_reference():
    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    class _Dataset:
        def __init__(self, *a, **k):
            pass

    mod("xarray", Dataset=_Dataset)
    mod("toolz", memoize=lambda f: f)
    mod("pendulum", now=lambda: None)

    class _Stream:
        def emit(self, *_):
            pass

    mod("streamz", Stream=_Stream)
    pkg = mod("triflow")
    pkg.__path__ = [os.path.join(REF, "triflow")]
    core = mod("triflow.core")
    core.__path__ = [os.path.join(REF, "triflow", "core")]
    plugins = mod("triflow.plugins")
    plugins.__path__ = []
    mod("triflow.plugins.container", TriflowContainer=object)
    out = {}
    for name in ["compilers", "routines", "fields", "model", "schemes", "simulation"]:
        spec = importlib.util.spec_from_file_location(
            "triflow.core." + name, os.path.join(REF, "triflow", "core", name + ".py"))
        m = importlib.util.module_from_spec(spec)
        sys.modules["triflow.core." + name] = m
        spec.loader.exec_module(m)
        setattr(core, name, m)
        out[name] = m
    # patch 1: one-argument Heaviside in the printed source
    from sympy.printing.numpy import NumPyPrinter
    NumPyPrinter._print_Heaviside = \
        lambda self, e: "Heaviside(%s)" % self._print(e.args[0])
    # ROW_general.__cache__ is memoized per (self, N) in the reference; our
    # no-op memoize keeps semantics (a fresh identity each call).
    return out


REFMODS = load_reference()
RefModel = REFMODS["model"].Model
ref_schemes = REFMODS["schemes"]


def ref_model(name):
    a = W.model_args(name)
    return RefModel(a["differential_equations"], a["dependent_variables"],
                    a["parameters"], a["help_functions"], compiler="numpy")


def ref_fields(model, x, **vars_):
    return RefFields(model._dep_vars, model._help_funcs, x=x, **vars_)


# ------------------------------------------------------------- expression dump
def dump_expressions():
    import sympy as sp
    from sympy import lambdify
    out = {}
    for name in W.MODELS:
        m = ref_model(name)
        table = {"amax": None, "amin": None, "Heaviside": None}
        fsrc = inspect.getsource(lambdify(m._symbolic_args, m.F_array.tolist(),
                                          modules=[table, "numpy"]))
        jsrc = inspect.getsource(lambdify(m._symbolic_args,
                                          m._J_sparse_array.tolist(),
                                          modules=[table, "numpy"]))
        out[name] = dict(
            F=[sp.srepr(e) for e in m.F_array.tolist()],
            J=[sp.srepr(e) for e in m.J_array.tolist()],
            sparse_indices=[int(i) for i in m._sparse_indices[0]],
            bounds=list(m._bounds), window_range=int(m._window_range),
            nvar=int(m._nvar), args=m._args, F_src=fsrc, J_src=jsrc)
    with open(os.path.join(HERE, "expr.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)


# --------------------------------------------------------------------- F and J
def _csc_triplet(J):
    J = J.tocsc()
    J.sum_duplicates()
    J.sort_indices()
    return J.indptr.astype(np.int64), J.indices.astype(np.int64), J.data


def fj_case(tag, name, x, fields, pars, seed=None):
    m = ref_model(name)
    f = ref_fields(m, x, **fields)
    F = m.F(f, pars)
    indptr, indices, data = _csc_triplet(m.J(f, pars))
    arrs = dict(x=x, F=F, J_indptr=indptr, J_indices=indices, J_data=data,
                periodic=np.array(bool(pars["periodic"])))
    for k, v in fields.items():
        arrs["field_" + k] = np.asarray(v, float)
    for k, v in pars.items():
        if k != "periodic":
            arrs["par_" + k] = np.asarray(v, float)
    np.savez_compressed(os.path.join(HERE, "fj_%s.npz" % tag), **arrs)
    print("fj", tag, F.shape, data.shape)


def dump_fj():
    rng = np.random.default_rng(7)
    for periodic in (True, False):
        p = "per" if periodic else "edge"
        c = W.readme(200)
        fj_case("advdiff_" + p, "advdiff", c["x"], c["fields"],
                dict(c["pars"], periodic=periodic))
        for acc in (1, 2, 3):
            c = W.burgers(512, acc)
            fj_case("burgers_up%d_%s" % (acc, p), c["model"], c["x"], c["fields"],
                    dict(c["pars"], periodic=periodic))
        c = W.kuramoto(512)
        fj_case("ks_" + p, "ks", c["x"], c["fields"], dict(periodic=periodic))
        c = W.film(256)
        fj_case("film_" + p, "film", c["x"], c["fields"],
                dict(c["pars"], periodic=periodic))
        x = np.linspace(0, 10, 100, endpoint=False)
        fj_case("helper_" + p, "helper", x,
                dict(U=np.cos(x * 2 * np.pi / 10), s=np.sin(x)),
                dict(k=1, periodic=periodic))
        fj_case("helperdx_" + p, "helper_dx", x,
                dict(U=np.cos(x * 2 * np.pi / 10), s=np.sin(x)),
                dict(k=.5, periodic=periodic))
        fj_case("upwindconst_" + p, "upwind_const", x,
                dict(U=np.cos(x * 2 * np.pi / 10), s=np.zeros_like(x)),
                dict(k=1, periodic=periodic))
        fj_case("coupled_" + p, "coupled", x,
                dict(U=np.cos(x * 2 * np.pi / 10), V=np.sin(x * 2 * np.pi / 10)),
                dict(k1=1., k2=.5, c1=.3, c2=-.2, periodic=periodic))
        fj_case("kdv_" + p, "kdv", x, dict(U=np.cos(x * 2 * np.pi / 10)),
                dict(c=1.5, b=.02, periodic=periodic))
        # per-node (array) parameter, random state
        fj_case("advdiff_arrpar_" + p, "advdiff", x,
                dict(U=rng.standard_normal(x.size)),
                dict(k=1e-2 * (1 + rng.random(x.size)), c=.3, periodic=periodic))
    # tiny grids (edge cases: N just above the stencil width)
    for N in (5, 6, 9):
        x = np.linspace(0, 1, N)
        fj_case("ks_tiny%d_per" % N, "ks", x, dict(U=rng.standard_normal(N)),
                dict(periodic=True))
        fj_case("ks_tiny%d_edge" % N, "ks", x, dict(U=rng.standard_normal(N)),
                dict(periodic=False))


# ---------------------------------------------------------------- trajectories
def _scheme(model, name, **kw):
    return getattr(ref_schemes, name)(model, **kw)


def run_fixed(model, scheme, x, fields, pars, dt, steps, hook=None, every=None):
    f = ref_fields(model, x, **fields)
    t = 0.0
    snaps = []
    hook = hook or ref_schemes.null_hook
    for i in range(steps):
        t, f = scheme(t, f, dt, pars, hook=hook)
        if every and (i + 1) % every == 0:
            snaps.append(f.uflat.copy())
    return t, f.uflat.copy(), np.array(snaps)


def dump_traj():
    out = {}
    # cfg 1: README, three parity modes + other schemes
    c = W.readme(200)
    m = ref_model("advdiff")
    for sname, kw in [("ROS3PRw", dict(time_stepping=False)), ("ROS2", {}),
                      ("Theta", dict(theta=1)), ("Theta05", dict(theta=.5)),
                      ("ROS3PRL", dict(time_stepping=False)),
                      ("RODASPR", dict(time_stepping=False))]:
        cls = "Theta" if sname.startswith("Theta") else sname
        t, U, snaps = run_fixed(m, _scheme(m, cls, **kw), c["x"], c["fields"],
                                c["pars"], c["dt"], 5, hook=W.readme_hook, every=1)
        out["readme_fixed_%s" % sname] = snaps
    # ROS3PRw own controller, called directly, tol=1e-1 + trace
    sch = _scheme(m, "ROS3PRw", tol=1e-1)
    trace = []
    orig = sch._fixed_step

    def traced(t, fields, dt, pars, hook=ref_schemes.null_hook):
        r = orig(t, fields, dt, pars, hook=hook)
        trace.append((t, dt, r[2]))
        return r
    sch._fixed_step = traced
    f = ref_fields(m, c["x"], **c["fields"])
    t = 0.0
    snaps, counts = [], []
    for i in range(5):
        n0 = len(trace)
        f, _ = W.readme_hook(t, f, c["pars"])          # Simulation's own hook call
        t, f = sch(t, f, c["dt"], c["pars"], hook=W.readme_hook)
        snaps.append(f.uflat.copy())
        counts.append(len(trace) - n0)
    out["readme_adaptive_ROS3PRw"] = np.array(snaps)
    out["readme_adaptive_trace"] = np.array(trace)
    out["readme_adaptive_counts"] = np.array(counts)
    # Simulation default (double wrapped)
    Sim = REFMODS["simulation"].Simulation

    class _Tmpl:
        def __init__(self, model):
            self.model = model

        def __call__(self, **kw):
            return ref_fields(self.model, kw.pop("x"), **kw)
    type(m).fields_template = property(lambda self: _Tmpl(self))
    sim = Sim(m, dict(x=c["x"], **c["fields"]), c["pars"], dt=c["dt"], tmax=c["tmax"],
              hook=W.readme_hook, scheme=ref_schemes.ROS3PRw)
    snaps = [fl.uflat.copy() for _, fl in sim]
    out["readme_simdefault_ROS3PRw"] = np.array(snaps)
    sim = Sim(m, dict(x=c["x"], **c["fields"]), c["pars"], dt=c["dt"], tmax=c["tmax"],
              hook=W.readme_hook, scheme=ref_schemes.ROS3PRw, time_stepping=False)
    out["readme_simfixed_ROS3PRw"] = np.array([fl.uflat.copy() for _, fl in sim])

    # cfg 2: Burgers upwind (acc 1, 2), N=2048, ROS2
    for acc in (1, 2):
        c = W.burgers(2048, acc)
        m = ref_model(c["model"])
        _, U, snaps = run_fixed(m, _scheme(m, "ROS2"), c["x"], c["fields"], c["pars"],
                                c["dt"], 50, every=10)
        out["burgers_up%d_2048" % acc] = snaps
    # cfg 3: KS, N=2048 and a non-multiple-of-256 N, ROS3PRw fixed
    for N in (2048, 1000):
        c = W.kuramoto(N)
        m = ref_model("ks")
        _, U, snaps = run_fixed(m, _scheme(m, "ROS3PRw", time_stepping=False), c["x"],
                                c["fields"], c["pars"], c["dt"], 50, every=10)
        out["ks_%d" % N] = snaps
    c = W.kuramoto(512)
    _, U, snaps = run_fixed(m, _scheme(m, "ROS3PRw", time_stepping=False), c["x"],
                            c["fields"], dict(periodic=False), c["dt"], 20, every=5)
    out["ks_512_edge"] = snaps
    # cfg 4: film, N=1024, Theta(1) and Theta(.5)
    for theta in (1, .5):
        c = W.film(1024, theta)
        m = ref_model("film")
        _, U, snaps = run_fixed(m, _scheme(m, "Theta", theta=theta), c["x"], c["fields"],
                                c["pars"], c["dt"], 100, every=20)
        out["film_1024_theta%g" % theta] = snaps
    # cfg 5: ensemble members, N=512 (fixture size) x 100 steps
    mem = W.ensemble_parity_subset(11)
    c = W.ensemble(512, mem)
    m = ref_model("advdiff")
    finals = []
    for idx in range(len(mem)):
        pars = dict(k=float(c["pars"]["k"][idx]), c=float(c["pars"]["c"][idx]),
                    periodic=False)
        _, U, _ = run_fixed(m, _scheme(m, "ROS3PRw", time_stepping=False), c["x"],
                            c["fields"], pars, c["dt"], 100, hook=W.readme_hook)
        finals.append(U)
    out["ensemble_512_members"] = mem
    out["ensemble_512_final"] = np.array(finals)
    # heat equation of the reference's own simulation tests (N=50)
    x = np.linspace(0, 10, 50, endpoint=False)
    T = np.cos(x * 2 * np.pi / 10)
    m = ref_model("heat")
    for sname, kw in [("ROS2", {}), ("ROS3PRw", dict(time_stepping=False)),
                      ("Theta", {})]:
        _, U, snaps = run_fixed(m, _scheme(m, sname, **kw), x, dict(T=T),
                                dict(k=1, periodic=True), 1.0, 20, every=5)
        out["heat50_%s" % sname] = snaps
    np.savez_compressed(os.path.join(HERE, "traj.npz"), **out)
    for k, v in out.items():
        print("traj", k, np.asarray(v).shape)


def _ens_member_full(job):
    idx, k, c_, N, steps, dt = job
    c = W.ensemble(N, [0])
    m = ref_model("advdiff")
    pars = dict(k=float(k), c=float(c_), periodic=False)
    _, U, _ = run_fixed(m, _scheme(m, "ROS3PRw", time_stepping=False), c["x"], c["fields"],
                        pars, dt, steps, hook=W.readme_hook)
    return idx, U


def dump_ensemble_full():
    """cfg 5 at its real size: the 64 parity members of SURVEY.md §8d (incl. 0, 127, 128,
    16384, 32767) at N = 4096 x 100 steps, each an independent run of the reference."""
    import multiprocessing as mp
    mem = W.ensemble_parity_subset(59)
    c = W.ensemble(4096, mem)
    jobs = [(i, c["pars"]["k"][i], c["pars"]["c"][i], 4096, 100, c["dt"]) for i in range(len(mem))]
    with mp.get_context("fork").Pool(min(8, os.cpu_count() or 1)) as pool:
        res = dict(pool.map(_ens_member_full, jobs))
    finals = np.array([res[i] for i in range(len(mem))])
    np.savez_compressed(os.path.join(HERE, "traj_ens4096.npz"), members=mem, final=finals)
    print("traj_ens4096", finals.shape)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "ens4096":
        dump_ensemble_full()
        sys.exit(0)
    dump_expressions()
    dump_fj()
    dump_traj()
    dump_ensemble_full()

"""Loads the REFERENCE's own ``triflow/core/*.py`` -- unmodified, by file path -- for tests
that drive the reference's objects (``Model``, ``Simulation``, ``F_Routine``) with this
repository's CUDA compiler plugin and schemes, and for the reference arm of ``bench.py``
(``--impl reference`` / ``cpu_baseline``), which times the reference's own CPU path.
Test / measurement infrastructure only: nothing under ``triflow_b200/`` imports it.

Source of the files: ``baseline/_ref/triflow/core`` (git-ignored copy made by
``__graft_entry__.build()`` from ``/root/reference``; it travels to the GPU box) or
``$TRIFLOW_REFERENCE``.  The import shims and the two compatibility patches are those of
SURVEY.md Appendix A / §8c (the same ones ``tests/golden/make_golden.py`` documents):
xarray / toolz / pendulum / streamz are not installed, SymPy 1.14 prints the two-argument
``Heaviside``, and ``BaseFields.uflat`` needs numpy < 1.23 -> a duck-typed Fields stand-in.
"""
import importlib.util
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CANDIDATES = [os.environ.get("TRIFLOW_REFERENCE", ""), os.path.join(ROOT, "baseline", "_ref"),
              "/root/reference"]
CORE_FILES = ["compilers", "routines", "fields", "model", "schemes", "simulation"]


def reference_root():
    for c in CANDIDATES:
        if c and os.path.exists(os.path.join(c, "triflow", "core", "model.py")):
            return c
    return None


class _RefArray(np.ndarray):
    @property
    def values(self):
        return self.view(np.ndarray)


class RefFields:
    """Duck-typed stand-in for the reference's xarray-based BaseFields (same uflat layout)."""

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


_loaded = None


def load_reference():
    """dict name -> module for compilers, routines, fields, model, schemes, simulation."""
    global _loaded
    if _loaded is not None:
        return _loaded
    ref = reference_root()
    if ref is None:
        raise FileNotFoundError("reference sources not found (baseline/_ref or /root/reference)")

    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    class _Dataset:
        def __init__(self, *a, **k):
            pass

    class _Stream:
        def emit(self, *_):
            pass

    mod("xarray", Dataset=_Dataset)
    mod("toolz", memoize=lambda f: f)
    mod("pendulum", now=lambda: None)
    mod("streamz", Stream=_Stream)
    pkg = mod("triflow")
    pkg.__path__ = [os.path.join(ref, "triflow")]
    core = mod("triflow.core")
    core.__path__ = [os.path.join(ref, "triflow", "core")]
    plugins = mod("triflow.plugins")
    plugins.__path__ = []
    mod("triflow.plugins.container", TriflowContainer=object)
    out = {}
    for name in CORE_FILES:
        spec = importlib.util.spec_from_file_location(
            "triflow.core." + name, os.path.join(ref, "triflow", "core", name + ".py"))
        m = importlib.util.module_from_spec(spec)
        sys.modules["triflow.core." + name] = m
        spec.loader.exec_module(m)
        setattr(core, name, m)
        out[name] = m
    from sympy.printing.numpy import NumPyPrinter
    NumPyPrinter._print_Heaviside = \
        lambda self, e: "Heaviside(%s)" % self._print(e.args[0])

    # Model.fields_template builds the xarray container: hand out the stand-in instead
    class _Tmpl:
        def __init__(self, model):
            self.model = model

        def __call__(self, **kw):
            return RefFields(self.model._dep_vars, self.model._help_funcs, x=kw.pop("x"), **kw)
    out["model"].Model.fields_template = property(lambda self: _Tmpl(self))
    _loaded = out
    return out

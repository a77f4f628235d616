"""triflow_b200: B200-native (sm_100a) implicit method-of-lines hot path behind
the triflow plugin API.  See DESIGN.md.

Public surface mirrors the reference package (``triflow/__init__.py:4-18``) for
the hot path only: ``Model``, ``Simulation``, ``schemes``.
"""

__version__ = "0.1.0"

_LAZY = {"Model": ("model", "Model"), "Simulation": ("simulation", "Simulation"),
         "schemes": ("schemes", None), "cuda_compiler": ("compiler", "cuda_compiler")}


def __getattr__(name):
    if name in _LAZY:
        import importlib
        modname, attr = _LAZY[name]
        mod = importlib.import_module("." + modname, __name__)
        return mod if attr is None else getattr(mod, attr)
    raise AttributeError(name)

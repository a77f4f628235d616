"""CPU oracle: numpy evaluation of F and J (TEST INFRASTRUCTURE, see oracle/__init__).

Restates the reference's ``numpy_compiler`` (``triflow/core/compilers.py:181-332``):

* ``lambdify`` of ``model.F_array`` / ``model._J_sparse_array`` over
  ``model._symbolic_args`` with the reference's function table
  (``compilers.py:196-219``) — ``Heaviside`` evaluates to 1 everywhere
  (``compilers.py:204-205``; SymPy >= 1.9 prints a two-argument Heaviside, which
  is accepted and ignored here, SURVEY.md §8c patch 2);
* ghost cells (``compilers.py:252-265``): periodic wrap or edge replication;
* shifted views ``U_m1, U, U_p1, ...`` (``compilers.py:266-277``);
* F interleaved node-major / variable-minor (``compilers.py:287-288``);
* J assembled as COO -> CSC with duplicate summation, through the same
  wrapped / clamped column table (``compilers.py:303-331``).
"""

import numpy as np
from scipy.sparse import csc_matrix
from sympy import lambdify


def _np_min(args):
    a, b = args
    return np.where(a < b, a, b)


def _np_max(args):
    a, b = args
    return np.where(a < b, b, a)


def _np_heaviside(a, h0=None):
    # the reference's table entry: where(a < 0, 1, 1)  (compilers.py:204-205)
    return np.where(np.asarray(a) < 0, 1, 1)


_TABLE = {"amax": _np_max, "amin": _np_min, "Heaviside": _np_heaviside}


def make_lambdas(model):
    f = lambdify(model._symbolic_args, expr=model.F_array.tolist(),
                 modules=[_TABLE, "numpy"])
    j = lambdify(model._symbolic_args, expr=model._J_sparse_array.tolist(),
                 modules=[_TABLE, "numpy"])
    return f, j


def _named_inputs(model, input_args):
    names = [*model._indep_vars, *model._dep_vars, *model._help_funcs,
             *model._pars, "periodic"]
    named = dict(zip(names, input_args))
    x = np.asarray(named["x"], dtype=np.float64)
    # the reference receives every parameter already broadcast to (N,)
    # (core/routines.py:40); do it here so that scalar parameters take the
    # same elementwise path.
    for p in model._pars:
        named[p] = np.asarray(named[p], dtype=np.float64) + x * 0
    return named, x


def _stencil_views(model, named, x):
    """Ghost-cell padding + shifted views (compilers.py:227-278)."""
    N = x.size
    dx = (x[-1] - x[0]) / (N - 1)
    periodic = named["periodic"]
    lo, hi = model._bounds
    half = int((model._window_range - 1) / 2)
    views = dict(named)
    views["dx"] = dx
    for name in model._symb_vars_with_spatial_diff_order:
        arr = np.asarray(named[name], dtype=np.float64)
        if periodic:
            ext = np.concatenate([arr[lo:], arr, arr[:hi]]) if lo else arr
        else:
            ext = np.concatenate([[arr[0]] * half, arr, [arr[-1]] * half])
        for o in range(lo, hi + 1):
            key = name if o == 0 else "%s_%s%d" % (name, "m" if o < 0 else "p", abs(o))
            views[key] = ext[o - lo: ext.size + o - hi]
    return views, N, half, periodic


def compute_F(model, f_func, *input_args):
    named, x = _named_inputs(model, input_args)
    views, N, _, _ = _stencil_views(model, named, x)
    rows = f_func(*[views[k] for k in model._args])
    rows = [np.broadcast_to(np.asarray(r, dtype=np.float64), (N,)) for r in rows]
    return np.stack(rows, axis=1).reshape(-1)


def compute_J(model, j_func, *input_args):
    named, x = _named_inputs(model, input_args)
    views, N, half, periodic = _stencil_views(model, named, x)
    nvar, w = model._nvar, model._window_range
    vals = j_func(*[views[k] for k in model._args])
    vals = np.stack([np.broadcast_to(np.asarray(v, dtype=np.float64), (N,))
                     for v in vals], axis=1)                     # (N, nnz)
    kk = np.asarray(model._sparse_indices[0])                    # flat J_array index
    eq = kk % nvar
    col = kk // nvar
    var = col % nvar
    off = col // nvar - half
    i = np.arange(N)[:, None]
    j = i + off[None, :]
    j = j % N if periodic else np.clip(j, 0, N - 1)
    rows = i * nvar + eq[None, :]
    cols = j * nvar + var[None, :]
    return csc_matrix((vals.reshape(-1), (rows.reshape(-1), cols.reshape(-1))),
                      shape=(N * nvar, N * nvar))


def numpy_compiler(model):
    """Plugin entry: ``Model(..., compiler=numpy_compiler)``."""
    f_func, j_func = make_lambdas(model)

    def F_function(*args):
        return compute_F(model, f_func, *args)

    def J_function(*args):
        return compute_J(model, j_func, *args)

    return F_function, J_function

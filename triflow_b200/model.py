"""Symbolic front-end: PDE strings -> finite-difference stencil F and Jacobian J.

This is the *input producer* of the hot path: a one-off host
computation.  It mirrors the attribute surface of the reference ``Model``
(reference ``triflow/core/model.py:138-311``) that a compiler plugin consumes:

    F_array, J_array, _J_sparse_array, _sparse_indices, _bounds,
    _window_range, _nvar, _args / _symbolic_args, _indep_vars, _dep_vars,
    _help_funcs, _pars, _symb_vars_with_spatial_diff_order

so that ``compiler(model) -> (F_function, J_function)`` plugins written for the
reference (``model.py:299-311``) work unchanged, and so that the expression
trees — and therefore the operation order the generated CUDA code reproduces —
are identical to the reference's.  ``tests/test_model_frontend.py`` pins the
expression text against strings dumped from the reference itself
(``tests/golden/make_golden.py``).

Differences from the reference, all deliberate:

* ``compiler`` defaults to ``"cuda"`` (the sm_100a compiler plugin of this
  package); ``"theano"``/``"numpy"`` are not shipped.  Any callable with the
  reference's plugin signature is accepted (the test-suite passes the CPU
  oracle this way).
* finite-difference tables are expressed as weight tables instead of
  per-order code.
"""

import logging
import pickle
import sys
from functools import partial

import numpy as np
import sympy as sp
from sympy import Derivative, Function, Max, Min, Symbol, SympifyError, sympify

from .fields import BaseFields
from .routines import F_Routine, J_Routine

log = logging.getLogger(__name__)
log.addHandler(logging.NullHandler())

# the reference raises the limit for deep expression trees (model.py:21)
sys.setrecursionlimit(max(sys.getrecursionlimit(), 40000))
EPS = 1e-6

# central differences: order -> [(offset, weight)] / dx**order (model.py:401-439)
_HALF = 1 / 2
_CENTRAL = {
    1: [(+1, _HALF), (-1, -_HALF)],
    2: [(+1, 1), (0, -2), (-1, 1)],
    3: [(-2, -_HALF), (-1, 1), (+1, -1), (+2, _HALF)],
    4: [(-2, 1), (-1, -4), (0, 6), (+1, -4), (+2, 1)],
}
# one-sided differences: accuracy -> (reach, backward weights, forward weights,
# denominator multiple of dx) (model.py:441-478)
_UPWIND = {
    1: (1, [(0, 1), (-1, -1)], [(+1, 1), (0, -1)], 1),
    2: (2, [(0, 3), (-1, -4), (-2, 1)], [(0, -3), (+1, 4), (+2, -1)], 2),
    3: (2, [(+1, 2), (0, 3), (-1, -6), (-2, 1)],
        [(-1, -2), (0, -3), (+1, 6), (+2, -1)], 6),
}


def _shifted(name, offset):
    if offset == 0:
        return Symbol(name)
    return Symbol("%s_%s%d" % (name, "m" if offset < 0 else "p", abs(offset)))


def _coerce(arg):
    if arg is None:
        return tuple()
    if isinstance(arg, str):
        return (arg,)
    return tuple(arg)


def _rebuild_model(eqs, deps, pars, helps, bdcs):
    return Model(eqs, deps, pars, helps, bdcs)


class Model:
    """Finite-difference approximation of ``dU/dt = F(U)`` and its Jacobian.

    Parameters follow the reference (``model.py:138-150``):
    ``Model(differential_equations, dependent_variables, parameters=None,
    help_functions=None, bdc_conditions=None, compiler=..., simplify=False,
    fdiff_jac=False, double=True, hold_compilation=False)``.
    """

    def __init__(self, differential_equations, dependent_variables,
                 parameters=None, help_functions=None, bdc_conditions=None,
                 compiler="cuda", simplify=False, fdiff_jac=False,
                 double=True, hold_compilation=False):
        self._double = double
        self._diff_eqs = _coerce(differential_equations)
        self._indep_vars = ("x",)
        self._dep_vars = _coerce(dependent_variables)
        self._pars = _coerce(parameters)
        self._help_funcs = _coerce(help_functions)
        self._bdcs = _coerce(bdc_conditions)
        self._nvar = len(self._dep_vars)

        x = Symbol("x")
        self._symb_indep_vars = (x,)
        self._symb_dep_vars = tuple(Function(n)(x) for n in self._dep_vars)
        self._symb_help_funcs = tuple(Function(n)(x) for n in self._help_funcs)
        self._symb_pars = sp.symbols(self._pars)
        if isinstance(self._symb_pars, Symbol):
            self._symb_pars = (self._symb_pars,)

        # stencil bookkeeping: name -> {(symbol, offset)} (model.py:220-223)
        self._symb_vars_with_spatial_diff_order = {
            name: {(Function(name), 0)}
            for name in self._dep_vars + self._help_funcs}

        self._symb_diff_eqs = self._parse(self._diff_eqs)
        self._symb_bdcs = self._parse(self._bdcs)
        approx = self._discretize(self._symb_diff_eqs)
        self._dbdcs = self._discretize(self._symb_bdcs)

        # half-widths of the stencil come from the dependent variables only
        # (model.py:244-247, 380-386)
        lo = hi = 0
        for name in self._dep_vars:
            offs = [o for _, o in self._symb_vars_with_spatial_diff_order[name]]
            lo, hi = min(lo, min(offs)), max(hi, max(offs))
        self._bounds = (lo, hi)
        self._window_range = hi - lo + 1

        # offset-major / variable-minor unknown ordering (model.py:252-262)
        offsets = range(lo, hi + 1)
        U = [_shifted(n, o) for o in offsets for n in self._dep_vars]
        self._discrete_variables = np.array(
            [_shifted(n, o) for o in offsets
             for n in self._dep_vars + self._help_funcs], dtype=object)

        self.F_array = np.array(approx)
        if simplify:
            self.F_array = np.array([eq.simplify() for eq in self.F_array.tolist()])
        if fdiff_jac:
            jac = [[(eq.subs(u, u + EPS) - eq) / EPS for u in U] for eq in approx]
        else:
            jac = [[eq.diff(u) for u in U] for eq in approx]
        # J_array[(col * nvar) + eq], col over U (model.py:279-281)
        self.J_array = np.array(jac).flatten("F")
        if simplify:
            self.J_array = np.array([e.expand().simplify()
                                     for e in self.J_array.tolist()])
        self._sparse_indices = np.where(self.J_array != 0)
        self._J_sparse_array = self.J_array[self._sparse_indices]

        if hold_compilation:
            return
        self.compile(compiler)

    # ------------------------------------------------------------------ parse
    def _namespace(self):
        """Names usable in equation strings (model.py:25-74)."""
        x = self._symb_indep_vars[0]

        def nth_derivative(order, expr):
            return Derivative(expr, x, order)

        ns = {"x": x}
        for order in range(1, 10):
            ns["d" + "x" * order] = partial(nth_derivative, order)
            for name in self._dep_vars + self._help_funcs:
                ns["d%s%s" % ("x" * order, name)] = Derivative(
                    Function(name)(x), x, order)
        return ns

    def _parse(self, equations):
        ns = self._namespace()
        x = self._symb_indep_vars[0]
        as_function = {Symbol(n): Function(n)(x) for n in self._dep_vars}
        try:
            return tuple(sympify(eq, locals=ns).xreplace(as_function).doit()
                         for eq in equations)
        except (TypeError, SympifyError):
            raise ValueError("badly formated differential equations")

    # ------------------------------------------------------------- discretize
    def _register(self, name, offsets):
        for o in offsets:
            if o != 0:
                self._symb_vars_with_spatial_diff_order[name].add(
                    (_shifted(name, o), o))

    def _finite_diff_scheme(self, U, order):
        name = str(U)
        if order not in _CENTRAL:
            raise NotImplementedError(
                "Finite difference up to 5th order not implemented yet")
        reach = 1 if order <= 2 else 2
        self._register(name, range(-reach, reach + 1))
        total = sum(w * _shifted(name, o) for o, w in _CENTRAL[order])
        dx = Symbol("dx")
        return total / dx if order == 1 else total / dx ** order

    def _upwind_scheme(self, a, U, accuracy):
        if accuracy not in _UPWIND:
            raise NotImplementedError("Upwind up to 2nd order not implemented yet")
        name = str(U)
        reach, backward, forward, mult = _UPWIND[accuracy]
        self._register(name, range(-reach, reach + 1))
        dx = Symbol("dx")
        denom = dx if mult == 1 else mult * dx
        Um = sum(w * _shifted(name, o) for o, w in backward) / denom
        Up = sum(w * _shifted(name, o) for o, w in forward) / denom
        return Max(a, 0) * Um + Min(a, 0) * Up

    def _discretize(self, equations):
        x = self._symb_indep_vars[0]
        every_field = self._symb_dep_vars + self._symb_help_funcs
        to_symbol = [(f, Symbol(str(f.func))) for f in every_field]
        out = []
        for eq in equations:
            for d in eq.find(Derivative):
                order = 0
                for wrt in d.args[1:]:
                    sym, n = (wrt, 1) if isinstance(wrt, Symbol) else (wrt[0], wrt[1])
                    if sym == x:
                        order = n
                var = Symbol(str(d.args[0].func))
                eq = eq.replace(d, self._finite_diff_scheme(var, order))
            eq = eq.subs(to_symbol)
            eq = eq.replace(Function("upwind"), self._upwind_scheme)
            out.append(eq.expand())
        return tuple(out)

    # ---------------------------------------------------------------- plug-in
    def compile(self, compiler):
        """Attach a compiler plugin (``model.py:299-311``)."""
        if compiler == "cuda":
            from .compiler import cuda_compiler as compiler
        elif isinstance(compiler, str):
            raise ValueError(
                "compiler %r is not shipped with triflow_b200: pass 'cuda' or a "
                "callable compiler(model) -> (F_function, J_function)" % compiler)
        F_function, J_function = compiler(self)
        names = self._dep_vars + self._help_funcs
        self.F = F_Routine(self.F_array, names, self._pars, F_function)
        self.J = J_Routine(self._J_sparse_array, names, self._pars, J_function)

    @property
    def fields_template(self):
        return BaseFields.factory1D(self._dep_vars, self._help_funcs)

    @property
    def _symbolic_args(self):
        return [*self._symb_indep_vars, *self._discrete_variables,
                *self._symb_pars, Symbol("dx")]

    @property
    def _args(self):
        return [str(a) for a in self._symbolic_args]

    # ------------------------------------------------------------ persistence
    def save(self, filename):
        with open(filename, "wb") as f:
            pickle.dump(self, f)

    @staticmethod
    def load(filename):
        with open(filename, "rb") as f:
            return pickle.load(f)

    def __reduce__(self):
        # like the reference (model.py:579-583) the compiler choice is not kept
        return (_rebuild_model, (self._diff_eqs, self._dep_vars, self._pars,
                                 self._help_funcs, self._bdcs))

    def __repr__(self):
        return ("%s\n\nVariables\n---------\nunknowns:       %s\n"
                "helpers:        %s\nparameters:     %s" % (
                    "\n".join(self._diff_eqs), ", ".join(self._dep_vars),
                    ", ".join(self._help_funcs) if self._pars else None,
                    ", ".join(self._pars) if self._pars else None))

"""F / J routine wrappers: the argument-marshalling boundary of the hot path.

Same call contract as the reference (``triflow/core/routines.py:8-91``): a
compiler plugin hands back ``ufunc(x, *dep_vars, *helpers, *pars, periodic)``
and these wrappers marshal ``fields`` / ``pars`` into that positional call.

One deliberate difference: the reference broadcasts every parameter to an
``(N,)`` vector before the call (``routines.py:40``); here a scalar parameter
is passed through as a 0-d float64 and only array parameters stay arrays, so a
compiler can tell "uniform" from "per-node" parameters.  Every arithmetic
operation the reference then performs is elementwise, so the values are
bit-identical either way; plugins that want the reference's vectors can
broadcast themselves (the CPU oracle does).
"""

import numpy as np
import sympy as sp


class ModelRoutine:
    def __init__(self, matrix, args, pars, ufunc, reduced=False):
        self.pars = list(pars) + ["periodic"]
        self.matrix = matrix
        self.args = args
        self._ufunc = ufunc

    def __repr__(self):
        return sp.Matrix(self.matrix.tolist()).__repr__()

    def _marshal(self, fields, pars):
        uargs = [np.asarray(fields["x"].values, dtype=np.float64),
                 *[np.asarray(fields[key].values, dtype=np.float64)
                   for key in self.args]]
        pargs = [np.asarray(pars[key], dtype=np.float64)
                 if key != "periodic" else pars[key]
                 for key in self.pars]
        return uargs, pargs


class F_Routine(ModelRoutine):
    """``F(fields, pars) -> ndarray (N*nvar,)`` (``routines.py:37-45``)."""

    def __call__(self, fields, pars):
        uargs, pargs = self._marshal(fields, pars)
        return self._ufunc(*uargs, *pargs)

    def diff_approx(self, fields, pars, eps=1e-3):
        """Brute-force finite-difference Jacobian (``routines.py:47-61``)."""
        U = fields.uflat
        J = np.zeros((U.size, U.size))
        F = self(fields, pars)
        for i in range(U.size):
            fields_plus = fields.copy()
            Up = fields_plus.uflat
            Up[i] += eps
            fields_plus.fill(Up)
            J[i] = (self(fields_plus, pars) - F) / eps
        return J.T


class J_Routine(ModelRoutine):
    """``J(fields, pars, sparse=True)`` -> ``csc_matrix`` or dense
    (``routines.py:82-91``)."""

    def __call__(self, fields, pars, sparse=True):
        uargs, pargs = self._marshal(fields, pars)
        J = self._ufunc(*uargs, *pargs)
        return J if sparse else J.todense()

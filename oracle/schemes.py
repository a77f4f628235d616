"""CPU oracle: implicit time schemes over scipy.sparse SuperLU
(TEST INFRASTRUCTURE, see oracle/__init__).

Restates the reference's ``triflow/core/schemes.py`` for the hot path:

* ``null_hook``                      <- ``schemes.py:29-30``
* ``time_stepping`` (Richardson)     <- ``schemes.py:33-66``
* ``ROW_general`` fixed / variable   <- ``schemes.py:69-238``
* ``ROS2, ROS3PRw, ROS3PRL, RODASPR``<- ``schemes.py:241-427`` (tableaux copied
  as numbers: they are the published constants of the methods)
* ``Theta``                          <- ``schemes.py:502-559``

The arithmetic order of every expression follows the reference so that results
are bit-identical given the same SciPy build (quirk kept on purpose: the
"predictor" error is ``||sum_i b_pred_i k_i||_inf`` because ``U_pred`` is built
from the already updated ``U``, ``schemes.py:164-174``).
"""

import numpy as np
import scipy.sparse as sps
import scipy.sparse.linalg  # noqa: F401  (sps.linalg)
from scipy.linalg import norm


def null_hook(t, fields, pars):
    return fields, pars


def time_stepping(scheme, tol=1e-1, ord=2, m=10, reject_factor=2):
    """Richardson step-doubling controller (``schemes.py:33-66``)."""
    state = {"dt": None}

    def one_step(t, fields, dt, pars, hook):
        dt_ = dt
        while True:
            _, coarse = scheme(t, fields, m * dt_, pars, hook)
            for _i in range(10):
                t, fields = scheme(t, fields, dt_, pars, hook)
            errs = [np.linalg.norm(coarse[key] - fields[key], ord) / (m ** 2 - 1)
                    for key in fields.dependent_variables]
            err = max(errs)
            dt_ = np.sqrt(dt ** 2 * tol / err)
            if dt_ < dt / reject_factor:
                continue
            break
        return t, fields, dt_

    def adaptive(t, fields, dt, pars, hook=null_hook):
        next_step = t + dt
        state["dt"] = state["dt"] if state["dt"] else dt
        while t + state["dt"] <= next_step:
            t, fields, state["dt"] = one_step(t, fields, state["dt"] / m, pars, hook)
        if t < next_step:
            t, fields = scheme(t, fields, next_step - t, pars, hook)
        return t, fields

    return adaptive


class ROW_general:
    """Rosenbrock-Wanner family (``schemes.py:69-238``)."""

    def __init__(self, model, alpha, gamma, b, b_pred=None, time_stepping=False,
                 tol=None, max_iter=None, dt_min=None, safety_factor=0.9,
                 recompute_target=True):
        self._model = model
        self._alpha, self._gamma = np.asarray(alpha, float), np.asarray(gamma, float)
        self._b, self._b_pred = b, b_pred
        self._s = len(b)
        self._time_control = time_stepping
        self._tol = tol
        self._safety_factor = safety_factor
        self._max_iter, self._dt_min = max_iter, dt_min
        self._recompute_target = recompute_target
        self._internal_dt = None
        self._internal_iter = None
        self._interp_cache = None
        self._eye = {}
        self.n_fixed_steps = 0          # oracle-side instrumentation only

    def _identity(self, n):
        if n not in self._eye:
            self._eye[n] = sps.eye(n, format="csc")
        return self._eye[n]

    def __call__(self, t, fields, dt, pars, hook=null_hook):
        if self._time_control:
            return self._variable_step(t, fields, dt, pars, hook=hook)
        t, fields, _ = self._fixed_step(t, fields, dt, pars, hook=hook)
        fields, pars = hook(t, fields, pars)
        return t, fields

    def _fixed_step(self, t, fields, dt, pars, hook=null_hook):
        self.n_fixed_steps += 1
        fields = fields.copy()
        fields, pars = hook(t, fields, pars)
        J = self._model.J(fields, pars)
        self._A = A = self._identity(fields.uflat.size) - self._gamma[0, 0] * dt * J
        luf = sps.linalg.factorized(A)
        ks = []
        stage = fields.copy()
        for i in range(self._s):
            stage.fill(fields.uflat
                       + sum([self._alpha[i, j] * ks[j] for j in range(i)]))
            F = self._model.F(stage, pars)
            ks.append(luf(dt * F
                          + dt * (J @ sum([self._gamma[i, j] * ks[j]
                                           for j in range(i)]) if i > 0 else 0)))
        U = fields.uflat.copy()
        U = U + sum([bi * ki for bi, ki in zip(self._b, ks)])
        U_pred = (U + sum([bi * ki for bi, ki in zip(self._b_pred, ks)])
                  if self._b_pred is not None else None)
        fields.fill(U)
        return t + dt, fields, (norm(U - U_pred, np.inf)
                                if U_pred is not None else None)

    def _variable_step(self, t, fields, dt, pars, hook=null_hook):
        self._next_time_step = t + dt
        self._internal_iter = 0
        if self._interp_cache is not None:
            try:
                fields.fill(self._interp_cache(self._next_time_step))
                return self._next_time_step, fields
            except (TypeError, ValueError):
                pass
        start = 1e-6 if self._internal_dt is None else self._internal_dt
        dt = self._internal_dt = min(start, dt) if self._recompute_target else start
        while True:
            self._err = None
            while self._err is None or self._err > self._tol:
                new_t, new_fields, self._err = self._fixed_step(t, fields, dt,
                                                                pars, hook)
                dt = self._internal_dt = (self._safety_factor * dt
                                          * np.sqrt(self._tol / self._err))
            if new_t >= self._next_time_step:
                if self._recompute_target:
                    t, fields, self._err = self._fixed_step(
                        t, fields, self._next_time_step - t, pars, hook)
                else:
                    from scipy.interpolate import interp1d
                    self._interp_cache = interp1d(
                        [t, new_t], [fields.uflat[None], new_fields.uflat[None]],
                        axis=0)
                    fields.fill(self._interp_cache(self._next_time_step))
                self._internal_iter += 1
                fields, pars = hook(t, fields, pars)
                return self._next_time_step, fields
            t = new_t
            fields = new_fields.copy()
            self._internal_iter += 1
            if self._internal_iter > (self._max_iter if self._max_iter
                                      else self._internal_iter + 1):
                raise RuntimeError("Rosebrock internal iteration "
                                   "above max iterations authorized")
            if dt < (self._dt_min if self._dt_min else dt * .5):
                raise RuntimeError("Rosebrock internal time step "
                                   "less than authorized")


# Tableaux: numeric constants of the published methods, as the reference
# spells them (schemes.py:250-256, 278-300, 326-353, 379-427).
def _tableau(s, gamma_diag, alpha, gamma):
    A = np.zeros((s, s))
    G = np.zeros((s, s))
    for (i, j), v in alpha.items():
        A[i, j] = v
    for (i, j), v in gamma.items():
        G[i, j] = v
    for i in range(s):
        G[i, i] = gamma_diag
    return A, G


TABLEAUX = {
    "ROS2": dict(
        s=2, gamma_diag=2.928932188134E-1,
        alpha={(1, 0): 1.0}, gamma={(1, 0): -5.857864376269E-1},
        b=[1 / 2, 1 / 2], b_pred=None),
    "ROS3PRw": dict(
        s=3, gamma_diag=7.8867513459481287e-01,
        alpha={(1, 0): 2.3660254037844388e+00, (2, 0): 5.0000000000000000e-01,
               (2, 1): 7.6794919243112270e-01},
        gamma={(1, 0): -2.3660254037844388e+00, (2, 0): -8.6791218280355165e-01,
               (2, 1): -8.7306695894642317e-01},
        b=[5.0544867840851759e-01, -1.1571687603637559e-01, 6.1026819762785800e-01],
        b_pred=[2.8973180237214197e-01, 1.0000000000000001e-01,
                6.1026819762785800e-01]),
    "ROS3PRL": dict(
        s=4, gamma_diag=4.3586652150845900e-01,
        alpha={(1, 0): .5, (2, 0): .5, (2, 1): .5, (3, 0): .5, (3, 1): .5, (3, 2): 0},
        gamma={(1, 0): -5.0000000000000000e-01, (2, 0): -7.9156480420464204e-01,
               (2, 1): 3.5244216792751432e-01, (3, 0): -4.9788969914518677e-01,
               (3, 1): 3.8607515441580453e-01, (3, 2): -3.2405197677907682e-01},
        b=[2.1103008548132443e-03, 8.8607515441580453e-01,
           -3.2405197677907682e-01, 4.3586652150845900e-01],
        b_pred=[5.0000000000000000e-01, 3.8752422953298199e-01,
                -2.0949226315045236e-01, 3.2196803361747034e-01]),
    "RODASPR": dict(
        s=6, gamma_diag=.25,
        alpha={(1, 0): 7.5E-1, (2, 0): 7.5162877593868457E-2,
               (2, 1): 2.4837122406131545E-2, (3, 0): 1.6532708886396510e0,
               (3, 1): 2.1545706385445562e-1, (3, 2): -1.3157488872766792e0,
               (4, 0): 1.9385003738039885e1, (4, 1): 1.2007117225835324e0,
               (4, 2): -1.9337924059522791e1, (4, 3): -2.4779140110062559e-1,
               (5, 0): -7.3844531665375115e0, (5, 1): -3.0593419030174646e-1,
               (5, 2): 7.8622074209377981e0, (5, 3): 5.7817993590145966e-1,
               (5, 4): 2.5e-1},
        gamma={(1, 0): -7.5e-1, (2, 0): -8.8644e-2, (2, 1): -2.868897e-2,
               (3, 0): -4.84700e0, (3, 1): -3.1583e-1, (3, 2): 4.9536568e0,
               (4, 0): -2.67694569e1, (4, 1): -1.5066459e0, (4, 2): 2.720013e1,
               (4, 3): 8.25971337e-1, (5, 0): 6.58762e0, (5, 1): 3.6807059e-1,
               (5, 2): -6.74235e0, (5, 3): -1.061963e-1, (5, 4): -3.57142857e-1},
        b=[-7.9683251690137014E-1, 6.2136401428192344E-2, 1.1198553514719862E00,
           4.7198362114404874e-1, -1.0714285714285714E-1, 2.5e-1],
        b_pred=[-7.3844531665375115e0, -3.0593419030174646e-1,
                7.8622074209377981e0, 5.7817993590145966e-1, 2.5e-1, 0]),
}


def _build(name):
    tab = TABLEAUX[name]
    alpha, gamma = _tableau(tab["s"], tab["gamma_diag"], tab["alpha"], tab["gamma"])
    return alpha, gamma, np.array(tab["b"]) if name == "ROS2" else tab["b"], tab["b_pred"]


class ROS2(ROW_general):
    def __init__(self, model):
        alpha, gamma, b, _ = _build("ROS2")
        super().__init__(model, alpha, gamma, b, time_stepping=False)


class _Adaptive(ROW_general):
    _name = None

    def __init__(self, model, tol=1e-1, time_stepping=True, max_iter=None,
                 dt_min=None, recompute_target=True):
        alpha, gamma, b, b_pred = _build(self._name)
        super().__init__(model, alpha, gamma, b, b_pred=b_pred,
                         time_stepping=time_stepping, tol=tol, max_iter=max_iter,
                         dt_min=dt_min, recompute_target=recompute_target)


class ROS3PRw(_Adaptive):
    _name = "ROS3PRw"


class ROS3PRL(_Adaptive):
    _name = "ROS3PRL"


class RODASPR(_Adaptive):
    _name = "RODASPR"


class Theta:
    """Theta scheme (``schemes.py:502-559``)."""

    def __init__(self, model, theta=1, solver=None):
        self._model = model
        self._theta = theta
        self._solver = solver if solver is not None else sps.linalg.spsolve

    def __call__(self, t, fields, dt, pars, hook=null_hook):
        fields = fields.copy()
        fields, pars = hook(t, fields, pars)
        F = self._model.F(fields, pars)
        J = self._model.J(fields, pars)
        U = fields.uflat
        B = dt * (F - self._theta * J @ U) + U
        A = sps.identity(U.size, format="csc") - self._theta * dt * J
        fields.fill(self._solver(A, B))
        fields, _ = hook(t + dt, fields, pars)
        return t + dt, fields

"""Implicit time schemes running on the device (drop-in for ``triflow.schemes``).

Same plugin contract as the reference (``triflow/core/schemes.py``;
``source_doc/source/user_guide.rst:296-342``): a scheme is built as
``scheme(model, **kwargs)`` and called as
``t, fields = scheme(t, fields, dt, pars, hook=hook)``; it returns a *new* fields
container and is stateful (``_internal_dt``).

What runs where:

* the whole step — J evaluation, ``A = I - gamma*dt*J``, banded LU, every stage
  (F evaluation, right-hand side, two triangular sweeps), the update and the
  embedded error estimate — is device code (``csrc/tf_kernels.cuh``) sequenced by
  ``libtriflow_b200.so``; the step-size controller of ``_variable_step``
  (``schemes.py:176-238``) runs in the library and reads one double per internal
  step;
* declarative hooks (:class:`Dirichlet`) are device kernels; any other Python
  hook is honoured through a host round trip per internal step (slow, exact);
* fields cross PCIe at the start and end of a ``__call__`` only.

The stage right-hand side uses the algebraically equivalent transformed
Rosenbrock form (``dt*J*w == (w - A*w)/gamma``), so ``J @ sum(gamma_ij k_j)`` of
``schemes.py:157-161`` needs no stored J and no matrix-vector product:
``k_i = A^-1 (dt*F(U_i) + sum_j c_ij k_j) - sum_j c_ij k_j``, ``c_ij =
gamma_ij/gamma_ii``.  ``Theta`` (``schemes.py:523-559``) is the one-stage case:
``B = dt*(F - theta*J@U) + U`` and ``A = I - theta*dt*J`` give
``U_new = U + A^-1 (dt*F(U))``.
"""

import ctypes as C

import numpy as np

from . import _lib
from .fields import LazyFields

# ------------------------------------------------------------------ tableaux
# Numeric constants of the published methods, spelled as the reference spells
# them (schemes.py:250-256, 278-300, 326-353, 379-427; truncated decimals kept).
_TABLEAUX = {
    "ROS2": (2, 2.928932188134E-1, {(1, 0): 1.0}, {(1, 0): -5.857864376269E-1},
             [1 / 2, 1 / 2], None),
    "ROS3PRw": (3, 7.8867513459481287e-01,
                {(1, 0): 2.3660254037844388e+00, (2, 0): 5.0000000000000000e-01,
                 (2, 1): 7.6794919243112270e-01},
                {(1, 0): -2.3660254037844388e+00, (2, 0): -8.6791218280355165e-01,
                 (2, 1): -8.7306695894642317e-01},
                [5.0544867840851759e-01, -1.1571687603637559e-01,
                 6.1026819762785800e-01],
                [2.8973180237214197e-01, 1.0000000000000001e-01,
                 6.1026819762785800e-01]),
    "ROS3PRL": (4, 4.3586652150845900e-01,
                {(1, 0): .5, (2, 0): .5, (2, 1): .5, (3, 0): .5, (3, 1): .5},
                {(1, 0): -5.0000000000000000e-01, (2, 0): -7.9156480420464204e-01,
                 (2, 1): 3.5244216792751432e-01, (3, 0): -4.9788969914518677e-01,
                 (3, 1): 3.8607515441580453e-01, (3, 2): -3.2405197677907682e-01},
                [2.1103008548132443e-03, 8.8607515441580453e-01,
                 -3.2405197677907682e-01, 4.3586652150845900e-01],
                [5.0000000000000000e-01, 3.8752422953298199e-01,
                 -2.0949226315045236e-01, 3.2196803361747034e-01]),
    "RODASPR": (6, .25,
                {(1, 0): 7.5E-1, (2, 0): 7.5162877593868457E-2,
                 (2, 1): 2.4837122406131545E-2, (3, 0): 1.6532708886396510e0,
                 (3, 1): 2.1545706385445562e-1, (3, 2): -1.3157488872766792e0,
                 (4, 0): 1.9385003738039885e1, (4, 1): 1.2007117225835324e0,
                 (4, 2): -1.9337924059522791e1, (4, 3): -2.4779140110062559e-1,
                 (5, 0): -7.3844531665375115e0, (5, 1): -3.0593419030174646e-1,
                 (5, 2): 7.8622074209377981e0, (5, 3): 5.7817993590145966e-1,
                 (5, 4): 2.5e-1},
                {(1, 0): -7.5e-1, (2, 0): -8.8644e-2, (2, 1): -2.868897e-2,
                 (3, 0): -4.84700e0, (3, 1): -3.1583e-1, (3, 2): 4.9536568e0,
                 (4, 0): -2.67694569e1, (4, 1): -1.5066459e0, (4, 2): 2.720013e1,
                 (4, 3): 8.25971337e-1, (5, 0): 6.58762e0, (5, 1): 3.6807059e-1,
                 (5, 2): -6.74235e0, (5, 3): -1.061963e-1, (5, 4): -3.57142857e-1},
                [-7.9683251690137014E-1, 6.2136401428192344E-2, 1.1198553514719862E00,
                 4.7198362114404874e-1, -1.0714285714285714E-1, 2.5e-1],
                [-7.3844531665375115e0, -3.0593419030174646e-1, 7.8622074209377981e0,
                 5.7817993590145966e-1, 2.5e-1, 0]),
}


def tableau(name):
    s, gdiag, a, g, b, bp = _TABLEAUX[name]
    alpha = np.zeros((s, s))
    gamma = np.zeros((s, s))
    for (i, j), v in a.items():
        alpha[i, j] = v
    for (i, j), v in g.items():
        gamma[i, j] = v
    gamma[np.arange(s), np.arange(s)] = gdiag
    return alpha, gamma, list(b), (None if bp is None else list(bp))


# ---------------------------------------------------------------------- hooks
def null_hook(t, fields, pars):
    return fields, pars


class Dirichlet:
    """Declarative boundary hook: ``Dirichlet(U=(1, 0))`` sets ``U[0]=1, U[-1]=0``
    wherever the reference would call ``hook(t, fields, pars)``; ``None`` leaves a
    side free.  Usable as a plain Python hook as well (``README.md:126-129``)."""

    def __init__(self, **values):
        self.values = {k: (v[0], v[1]) for k, v in values.items()}

    def __call__(self, t, fields, pars):
        if isinstance(fields, LazyFields) and fields._real is None:
            return fields, pars      # still on the device: the scheme applies the hook there
        for var, (left, right) in self.values.items():
            if left is not None:
                fields[var][0] = left
            if right is not None:
                fields[var][-1] = right
        return fields, pars


# ------------------------------------------------- Richardson wrapper (host side)
def time_stepping(scheme, tol=1e-1, ord=2, m=10, reject_factor=2):
    """Richardson controller the reference's ``Simulation`` wraps around every
    scheme (reference ``schemes.py:33-66``, ``simulation.py:190-197``): one
    coarse ``m*dt`` step against ten fine ``dt`` steps."""
    internal = [None]

    def attempt(t, fields, dt, pars, hook):
        trial = dt
        while True:
            _, coarse = scheme(t, fields, m * trial, pars, hook)
            for _ in range(10):
                t, fields = scheme(t, fields, trial, pars, hook)
            err = max(np.linalg.norm(coarse[k] - fields[k], ord) / (m ** 2 - 1)
                      for k in fields.dependent_variables)
            trial = np.sqrt(dt ** 2 * tol / err)
            if trial < dt / reject_factor:
                continue
            return t, fields, trial

    def adaptive(t, fields, dt, pars, hook=null_hook):
        target = t + dt
        if not internal[0]:
            internal[0] = dt
        while t + internal[0] <= target:
            t, fields, internal[0] = attempt(t, fields, internal[0] / m, pars, hook)
        if t < target:
            t, fields = scheme(t, fields, target - t, pars, hook)
        return t, fields

    return adaptive


# -------------------------------------------------------------- device binding
class _DeviceScheme:
    """Tableau on the library side + a device state per (grid, boundary, variant)."""

    def __init__(self, model, alpha, gamma, b, b_pred):
        cm = getattr(model, "_cuda", None)
        if cm is None:
            raise TypeError(
                "triflow_b200 schemes need a model compiled with the CUDA compiler "
                "(Model(..., compiler='cuda')); there is no CPU path")
        self._cm = cm
        self._model = model
        self._alpha = np.ascontiguousarray(alpha, dtype=np.float64)
        self._gamma = np.ascontiguousarray(gamma, dtype=np.float64)
        self._b = np.ascontiguousarray(b, dtype=np.float64)
        self._b_pred = (None if b_pred is None
                        else np.ascontiguousarray(b_pred, dtype=np.float64))
        self._s = len(self._b)
        self._handle = None
        self._state = None
        self._state_key = None
        self._generation = 0          # bumped whenever the device state changes
        self.lazy = False             # opt-in: return LazyFields, keep U on the device

    @property
    def handle(self):
        if self._handle is None:
            h = C.c_void_p()
            _lib.check(_lib.lib().tf_scheme_create(
                self._cm.ctx, self._s, _lib.dptr(self._alpha), _lib.dptr(self._gamma),
                _lib.dptr(self._b), _lib.dptr(self._b_pred), C.byref(h)))
            self._handle = h
        return self._handle

    def _bind(self, fields, pars):
        """Device state for this grid, with grid / helpers / parameters / unknowns
        uploaded."""
        x = np.asarray(fields["x"].values, dtype=np.float64)
        N = x.size
        periodic = bool(pars["periodic"])
        node_pars = self._cm.node_pars_of(pars, N, 1)
        key = (N, periodic, node_pars)
        if key != self._state_key:
            if self._state is not None:
                self._state.close()
            self._state = self._cm.new_state(pars, N, 1, periodic)
            self._state_key = key
        st = self._state
        named = {k: fields[k].values for k in self._model._help_funcs}
        st.set_inputs(x, named, pars)
        if not (isinstance(fields, LazyFields) and key == getattr(self, "_resident_key", None)
                and fields.is_resident(self)):
            st.upload(u=fields.uflat.reshape(1, -1))
        self._resident_key = key
        self._generation += 1         # whatever happens next changes the device state
        return st

    def _set_hook(self, st, hook):
        lib = _lib.lib()
        _lib.check(lib.tf_hook_clear(st.h))
        if isinstance(hook, Dirichlet):
            for var, (left, right) in hook.values.items():
                e = list(self._model._dep_vars).index(var)
                _lib.check(lib.tf_hook_set_dirichlet(
                    st.h, e, left is not None, 0.0 if left is None else float(left),
                    right is not None, 0.0 if right is None else float(right)))

    @staticmethod
    def _on_device(hook):
        return hook is null_hook or isinstance(hook, Dirichlet) or \
            getattr(hook, "__name__", "") == "null_hook"

    def _result(self, fields, st):
        if self.lazy:
            template = fields._template if isinstance(fields, LazyFields) else fields
            return LazyFields(template, self, self._generation)
        out = fields.copy()
        out.fill(st.download()[0])
        return out

    # -- one _fixed_step with an arbitrary Python hook (host round trip)
    def _fixed_step_host_hook(self, t, fields, dt, pars, hook):
        fields = fields.copy()
        fields, pars = hook(t, fields, pars)
        st = self._bind(fields, pars)
        self._set_hook(st, null_hook)
        err = np.empty(1)
        _lib.check(_lib.lib().tf_scheme_step(st.h, self.handle, float(dt), 1, _lib.dptr(err)))
        return t + dt, self._result(fields, st), (float(err[0]) if self._b_pred is not None
                                                  else None), pars

    def run_fixed(self, t, fields, dt, n_steps, pars, hook=null_hook):
        """``n_steps`` fixed steps without leaving the device (throughput path)."""
        if not self._on_device(hook):
            for _ in range(n_steps):
                t, fields, _, pars = self._fixed_step_host_hook(t, fields, dt, pars, hook)
                fields, pars = hook(t, fields, pars)
            return t, fields
        if not isinstance(fields, LazyFields):
            fields, pars = hook(t, fields.copy(), pars)
        st = self._bind(fields, pars)
        self._set_hook(st, hook)
        _lib.check(_lib.lib().tf_scheme_step(st.h, self.handle, float(dt), int(n_steps), None))
        return t + n_steps * dt, self._result(fields, st)


class ROW_general(_DeviceScheme):
    """Rosenbrock-Wanner family (reference ``schemes.py:69-238``)."""

    def __init__(self, model, alpha, gamma, b, b_pred=None, time_stepping=False,
                 tol=None, max_iter=None, dt_min=None, safety_factor=0.9,
                 recompute_target=True):
        super().__init__(model, alpha, gamma, b, b_pred)
        self._internal_dt = None
        self._time_control = time_stepping
        self._internal_iter = None
        self._tol = tol
        self._safety_factor = safety_factor
        self._max_iter = max_iter
        self._dt_min = dt_min
        self._recompute_target = recompute_target
        self._interp_cache = None
        self._err = None
        self.n_fixed_steps = 0

    def __call__(self, t, fields, dt, pars, hook=null_hook):
        if self._time_control:
            if self._b_pred is None:
                raise NotImplementedError("time stepping needs the b predictor coefficients")
            if self._tol is None:
                raise ValueError("time_stepping=True needs a tolerance")
            return self._variable_step(t, fields, dt, pars, hook)
        if self._on_device(hook):
            st = self._bind(fields, pars)
            self._set_hook(st, hook)
            err = np.empty(1)
            _lib.check(_lib.lib().tf_scheme_step(st.h, self.handle, float(dt), 1,
                                                 _lib.dptr(err)))
            self._err = float(err[0])
            self.n_fixed_steps += 1
            return t + dt, self._result(fields, st)
        t, fields, self._err, pars = self._fixed_step_host_hook(t, fields, dt, pars, hook)
        self.n_fixed_steps += 1
        fields, pars = hook(t, fields, pars)
        return t, fields

    def _variable_step(self, t, fields, dt, pars, hook):
        if self._recompute_target and self._on_device(hook):
            st = self._bind(fields, pars)
            self._set_hook(st, hook)
            idt = C.c_double(-1.0 if self._internal_dt is None else self._internal_dt)
            nfs = C.c_int(0)
            err = C.c_double(0.0)
            rc = _lib.lib().tf_scheme_advance(
                st.h, self.handle, float(t), float(dt), float(self._tol),
                float(self._safety_factor), int(self._max_iter or 0),
                float(self._dt_min or 0.0), 1, C.byref(idt), C.byref(nfs), C.byref(err))
            self._internal_dt = idt.value
            self.n_fixed_steps += nfs.value
            _lib.check(rc)
            self._err = err.value
            return t + dt, self._result(fields, st)
        # arbitrary Python hook, or recompute_target=False (output interpolated between the
        # internal steps, schemes.py:183-187,217-222): the controller of schemes.py:176-238 on
        # the host, every _fixed_step on the device.
        next_t = t + dt
        self._internal_iter = 0
        if not self._recompute_target:
            try:                                      # target inside the last internal step?
                fields = fields.copy()
                fields.fill(np.asarray(self._interp_cache(next_t)).ravel())
                return next_t, fields
            except (TypeError, ValueError):           # no cache yet / target beyond it
                pass
            dt = self._internal_dt = 1e-6 if self._internal_dt is None else self._internal_dt
        else:
            dt = self._internal_dt = min(1e-6 if self._internal_dt is None
                                         else self._internal_dt, dt)
        while True:
            self._err = None
            while self._err is None or self._err > self._tol:
                new_t, new_fields, self._err, _ = self._fixed_step_host_hook(
                    t, fields, dt, pars, hook)
                self.n_fixed_steps += 1
                dt = self._internal_dt = (self._safety_factor * dt
                                          * np.sqrt(self._tol / self._err))
            if new_t >= next_t:
                if self._recompute_target:
                    t, fields, self._err, _ = self._fixed_step_host_hook(
                        t, fields, next_t - t, pars, hook)
                    self.n_fixed_steps += 1
                else:
                    from scipy.interpolate import interp1d
                    self._interp_cache = interp1d([t, new_t], [fields.uflat[None],
                                                               new_fields.uflat[None]], axis=0)
                    fields = fields.copy()
                    fields.fill(np.asarray(self._interp_cache(next_t)).ravel())
                self._internal_iter += 1
                fields, pars = hook(t, fields, pars)
                return next_t, fields
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


class ROS2(ROW_general):
    """Second order Rosenbrock scheme, fixed step (``schemes.py:241-256``)."""

    def __init__(self, model, lazy=False):
        alpha, gamma, b, _ = tableau("ROS2")
        super().__init__(model, alpha, gamma, b, time_stepping=False)
        self.lazy = lazy


class _Embedded(ROW_general):
    _name = None

    def __init__(self, model, tol=1e-1, time_stepping=True, max_iter=None, dt_min=None,
                 recompute_target=True, lazy=False):
        alpha, gamma, b, b_pred = tableau(self._name)
        super().__init__(model, alpha, gamma, b, b_pred=b_pred,
                         time_stepping=time_stepping, tol=tol, max_iter=max_iter,
                         dt_min=dt_min, recompute_target=recompute_target)
        self.lazy = lazy


class ROS3PRw(_Embedded):
    """Third order Rosenbrock-W scheme with embedded error control
    (``schemes.py:259-300``)."""
    _name = "ROS3PRw"


class ROS3PRL(_Embedded):
    """``schemes.py:303-353``."""
    _name = "ROS3PRL"


class RODASPR(_Embedded):
    """``schemes.py:356-427`` (the reference ``Simulation`` default)."""
    _name = "RODASPR"


class Theta(_DeviceScheme):
    """Theta scheme (``schemes.py:502-559``): theta=1 backward Euler, 0.5
    Crank-Nicolson.  By default the solve is the device banded solver; a ``solver(A, b)``
    given by the user is called on the host like the reference does (F and J still come from
    the device)."""

    def __init__(self, model, theta=1, solver=None, lazy=False):
        if theta == 0 and solver is None:
            raise ValueError("theta=0 (explicit Euler) has no implicit system to solve")
        self._theta = theta
        self._solver = solver
        super().__init__(model, np.zeros((1, 1)), np.array([[float(theta)]]), [1.0], None)
        self.lazy = lazy

    def __call__(self, t, fields, dt, pars, hook=null_hook):
        if self._solver is not None:
            # a user-supplied solver(A, b) (schemes.py:518-521): the reference's own sequence
            # (:548-559) with F and J evaluated on the device, A handed over as scipy CSC
            import scipy.sparse as sps
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
        if self._on_device(hook):
            st = self._bind(fields, pars)
            self._set_hook(st, hook)
            _lib.check(_lib.lib().tf_scheme_step(st.h, self.handle, float(dt), 1, None))
            return t + dt, self._result(fields, st)
        t, fields, _, pars = self._fixed_step_host_hook(t, fields, dt, pars, hook)
        fields, _ = hook(t, fields, pars)
        return t, fields

"""Parameter-sweep ensembles: many independent systems stepped together.

The reference has no batch API (one ``Model`` + scheme run per parameter set,
SURVEY.md §8d cfg 5); here ``batch`` systems share a grid and a model and differ
in initial state and/or parameter values.  Each system is an independent banded
problem (one CTA tile chain per system), so an ensemble shards across GPUs with
no data-path communication (:mod:`triflow_b200.distributed`).
"""

import numpy as np

from . import _lib
from .schemes import Dirichlet, null_hook


class Ensemble:
    """``batch`` systems of one model on one device.

    ``fields``: dict name -> ``(N,)`` (shared) or ``(batch, N)``.
    ``pars``: dict name -> scalar, ``(batch,)`` (one value per member), ``(N,)``
    or ``(batch, N)`` (per-node arrays), plus ``periodic``.
    ``scheme``: a scheme instance from :mod:`triflow_b200.schemes` (its tableau is
    used; stepping is fixed-step).
    """

    def __init__(self, model, scheme, x, fields, pars, hook=null_hook, batch=None,
                 reuse_constant_factor=False, ctx=None):
        self.model, self.scheme = model, scheme
        cm = model._cuda
        x = np.asarray(x, dtype=np.float64)
        N = x.size
        if batch is None:
            batch = 1
            for v in list(fields.values()) + [pars[p] for p in model._pars]:
                if np.ndim(v) == 2 or (np.ndim(v) == 1 and np.shape(v)[0] != N):
                    batch = max(batch, np.shape(v)[0])
        self.batch, self.N = int(batch), N
        from .compiler import value_kind
        for name, v in fields.items():                         # raises on ambiguous / bad shapes
            value_kind(np.shape(v), self.batch, N, name, field=True)
        for name in model._pars:
            value_kind(np.shape(pars[name]), self.batch, N, name)
        self.nvar = model._nvar
        self.pars = dict(pars)
        self.state = cm.new_state(pars, N, self.batch, bool(pars["periodic"]), ctx=ctx)
        named = {h: fields[h] for h in model._help_funcs}
        self.state.set_inputs(x, named, pars)
        u = np.stack([np.broadcast_to(np.asarray(fields[v], dtype=np.float64),
                                      (self.batch, N)) for v in model._dep_vars], axis=2)
        self.state.upload(u=u.reshape(self.batch, N * self.nvar))
        if not scheme._on_device(hook):
            raise ValueError("ensembles take declarative hooks (schemes.Dirichlet) only")
        scheme._set_hook(self.state, hook)
        self.factor_reuse = False
        if reuse_constant_factor:
            # legitimate only when J does not depend on the state (SURVEY.md §7
            # "State-independent J"); the reference refactorises every step anyway
            if not self.state.variant.lowered.jacobian_is_constant:
                raise ValueError("reuse_constant_factor needs a state-independent Jacobian")
            _lib.check(_lib.lib().tf_state_set_factor_reuse(self.state.h, 1))
            self.factor_reuse = True
        self.t = 0.0

    def set_fusion(self, enable):
        """Which step kernels may be used: ``True`` / ``1`` automatic (system-resident kernel
        when a system fits one CTA, grid-resident single-launch step where it pays), ``False``
        / ``0`` the per-kernel pipeline only, ``"rt"`` the run-time-stage variant of the
        system-resident kernel, ``"grid"`` the grid-resident step wherever it applies.  All
        paths run the same algorithm and agree to rounding."""
        mode = {"rt": 2, "grid": 3}.get(enable, int(bool(enable)))
        _lib.check(_lib.lib().tf_state_set_fusion(self.state.h, mode))

    def upload(self, u):
        """``u``: (batch, N*nvar) in uflat layout."""
        self.state.upload(u=np.asarray(u).reshape(self.batch, self.N * self.nvar))

    def step(self, dt, n_steps=1, want_err=False):
        err = np.empty(self.batch) if want_err else None
        _lib.check(_lib.lib().tf_scheme_step(self.state.h, self.scheme.handle, float(dt),
                                             int(n_steps), _lib.dptr(err)))
        self.t += n_steps * dt
        return err

    def advance(self, dt, tol=None, safety_factor=None, max_iter=None, dt_min=None):
        """Advance every member to ``t + dt`` with its own embedded-error step-size
        controller (reference ``ROW_general._variable_step``, ``core/schemes.py:176-238``,
        run per member on the device).  Defaults come from the scheme instance.
        Returns the number of internal ``_fixed_step`` evaluations per member."""
        import ctypes as C
        sch = self.scheme
        if getattr(sch, "_b_pred", None) is None:
            raise NotImplementedError("time stepping needs the b predictor coefficients")
        tol = sch._tol if tol is None else tol
        if tol is None:
            raise ValueError("adaptive stepping needs a tolerance")
        safety = sch._safety_factor if safety_factor is None else safety_factor
        max_iter = sch._max_iter if max_iter is None else max_iter
        dt_min = sch._dt_min if dt_min is None else dt_min
        if not hasattr(self, "_internal_dt"):
            self._internal_dt = np.full(self.batch, -1.0)
        nfs = (C.c_int * self.batch)()
        fail = (C.c_int * self.batch)()
        rc = _lib.lib().tf_ensemble_advance(
            self.state.h, sch.handle, float(self.t), float(dt), float(tol), float(safety),
            int(max_iter or 0), float(dt_min or 0.0), _lib.dptr(self._internal_dt), nfs, fail)
        self.n_fixed_steps = np.array(nfs[:])
        self.failed = np.array(fail[:])
        _lib.check(rc)
        self.t += dt
        return self.n_fixed_steps

    def advance_simulation_default(self, dt, tol=1e-1, m=10, reject_factor=2):
        """Advance every member to ``t + dt`` the way an un-flagged ``Simulation(...)`` of
        the reference does: its Richardson controller ``schemes.time_stepping``
        (``core/schemes.py:33-66``, applied by ``core/simulation.py:190-197`` to every scheme)
        around the scheme -- which, for the Rosenbrock schemes constructed with a tolerance
        (``time_stepping=True``), runs its own embedded-error controller inside every call.
        Each member has its own controllers; everything runs on the device.
        Returns ``(scheme calls, fixed steps)`` per member."""
        import ctypes as C
        sch = self.scheme
        inner = bool(getattr(sch, "_time_control", False))
        if inner and (getattr(sch, "_b_pred", None) is None or sch._tol is None):
            raise ValueError("the wrapped scheme's own controller needs b_pred and a tolerance")
        if not hasattr(self, "_outer_dt"):
            self._outer_dt = np.zeros(self.batch)            # None
        if not hasattr(self, "_internal_dt"):
            self._internal_dt = np.full(self.batch, -1.0)
        calls, nfs, fail = ((C.c_int * self.batch)() for _ in range(3))
        rc = _lib.lib().tf_ensemble_richardson(
            self.state.h, sch.handle, int(inner), float(self.t), float(dt), float(tol), int(m),
            float(reject_factor), float(sch._tol or 0.0) if inner else 0.0,
            float(getattr(sch, "_safety_factor", 0.9)), int(getattr(sch, "_max_iter", 0) or 0),
            float(getattr(sch, "_dt_min", 0.0) or 0.0), _lib.dptr(self._outer_dt),
            _lib.dptr(self._internal_dt), calls, nfs, fail)
        self.n_scheme_calls = np.array(calls[:])
        self.n_fixed_steps = np.array(nfs[:])
        self.failed = np.array(fail[:])
        _lib.check(rc)
        self.t += dt
        return self.n_scheme_calls, self.n_fixed_steps

    def download(self, out=None):
        return self.state.download(out)

    def download_to_torch(self):
        """The states as a ``torch`` CUDA tensor ``(batch, N*nvar)`` on this rank's GPU (no host
        copy): the send buffer of :func:`triflow_b200.distributed.gather_members`."""
        import ctypes as C
        import torch
        dev = torch.device("cuda", torch.cuda.current_device())
        out = torch.empty((self.batch, self.N * self.nvar), dtype=torch.float64, device=dev)
        _lib.check(_lib.lib().tf_state_download_device(self.state.h, C.c_void_p(out.data_ptr())))
        _lib.check(_lib.lib().tf_ctx_sync(self.state.ctx))
        return out

    def member_fields(self, r, u=None):
        u = self.download() if u is None else u
        cols = u[r].reshape(self.N, self.nvar)
        return {v: cols[:, e].copy() for e, v in enumerate(self.model._dep_vars)}

    def sync(self):
        _lib.check(_lib.lib().tf_ctx_sync(self.state.ctx))


def _members(v, lo, hi, batch, N, field=False):
    """Slice the member axis of a field / parameter value (same shapes as Ensemble)."""
    from .compiler import value_kind
    if value_kind(np.shape(v), batch, N, field=field) in ("member", "member_node"):
        return np.asarray(v)[lo:hi]
    return v


class HostPipeline:
    """Ensemble stepped from / to host buffers: ``upload -> step -> download``.

    This is the reference's calling convention (``scheme(t, fields, dt, pars)`` takes and
    returns host fields, ``core/schemes.py:102-135``) for a whole ensemble.  The members
    are split into ``groups`` contiguous blocks, each on its own context (stream) in
    asynchronous mode, so the PCIe upload of one block, the stepping of another and the
    download of a third proceed together; with pinned buffers (``_lib.pinned_empty``) a
    call costs about ``max(upload, download)`` instead of their sum plus the step.
    """

    def __init__(self, model, scheme, x, fields, pars, hook=null_hook, batch=None, groups=8):
        x = np.asarray(x, dtype=np.float64)
        N = x.size
        if batch is None:
            batch = 1
            for v in list(fields.values()) + [pars[p] for p in model._pars]:
                if np.ndim(v) == 2 or (np.ndim(v) == 1 and np.shape(v)[0] != N):
                    batch = max(batch, np.shape(v)[0])
        self.batch, self.N, self.nvar = int(batch), N, model._nvar
        groups = max(1, min(int(groups), self.batch))
        edges = [self.batch * g // groups for g in range(groups + 1)]
        self.ranges = [(lo, hi) for lo, hi in zip(edges[:-1], edges[1:]) if hi > lo]
        lib = _lib.lib()
        self.parts = []
        for lo, hi in self.ranges:
            ctx = _lib.new_context()
            f = {k: _members(v, lo, hi, self.batch, N, field=True) for k, v in fields.items()}
            p = {k: (_members(v, lo, hi, self.batch, N) if k != "periodic" else v)
                 for k, v in pars.items()}
            part = Ensemble(model, scheme, x, f, p, hook=hook, batch=hi - lo, ctx=ctx)
            part.sync()
            _lib.check(lib.tf_ctx_set_async(ctx, 1))
            self.parts.append(part)
        self.scheme = scheme
        self.t = 0.0

    def step_host(self, u_in, u_out, dt, n_steps=1):
        """``u_in`` -> ``n_steps`` fixed steps -> ``u_out``; both ``(batch, N*nvar)`` in the
        ``uflat`` layout, C-contiguous float64 (pinned for overlap)."""
        lib = _lib.lib()
        width = self.N * self.nvar
        u_in = np.asarray(u_in).reshape(self.batch, width)
        u_out = np.asarray(u_out).reshape(self.batch, width)
        for (lo, hi), part in zip(self.ranges, self.parts):
            _lib.check(lib.tf_state_upload(part.state.h, None, _lib.dptr(u_in[lo:hi]), None,
                                           None, None))
            _lib.check(lib.tf_scheme_step(part.state.h, self.scheme.handle, float(dt),
                                          int(n_steps), None))
            _lib.check(lib.tf_state_download(part.state.h, _lib.dptr(u_out[lo:hi])))
        self.sync()
        self.t += n_steps * dt
        return u_out

    def sync(self):
        """Wait for every block; raises if a factorisation failed (as the blocking API)."""
        for part in self.parts:
            part.sync()
        for (lo, hi), part in zip(self.ranges, self.parts):
            bad = np.flatnonzero(part.state.status())
            if bad.size:
                raise RuntimeError("banded factorisation failed for system %d (status %d)"
                                   % (lo + bad[0], part.state.status()[bad[0]]))

    def launch_count(self):
        lib = _lib.lib()
        return sum(lib.tf_ctx_launch_count(p.state.ctx) for p in self.parts)

    def close(self):
        for part in self.parts:
            part.sync()
            part.state.close()
            _lib.lib().tf_ctx_destroy(part.state.ctx)
        self.parts = []


__all__ = ["Ensemble", "HostPipeline", "Dirichlet"]

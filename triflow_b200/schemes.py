"""placeholder (filled in below)"""
import numpy as np


def null_hook(t, fields, pars):
    return fields, pars


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

"""Host driver that iterates a scheme over output times (caller of the hot path).

Mirror of the part of the reference ``Simulation`` that the hot path is called
from (reference ``triflow/core/simulation.py:160-261``): hook before each
output step, last ``dt`` clipped to ``tmax``, ``scheme(t, fields, dt, pars,
hook=hook)``, CPU timers, post-processes, stop at ``isclose(t, tmax)``.  The
streamz / container / display plumbing of the reference (``simulation.py:
352-438``) is out of scope; ``stream`` is a minimal
synchronous emitter so user code that does ``simul.stream.sink(f)`` still runs.

Quirk kept for drop-in behaviour (``simulation.py:190-197``): with the default
``time_stepping=True`` *every* scheme instance — ROS3PRw included — is wrapped
by the Richardson controller, because the reference compares an instance with
classes.  Pass ``time_stepping=False`` for fixed-step runs.
"""

import inspect
import time
import warnings
from collections import namedtuple
from uuid import uuid1

from numpy import isclose

from . import schemes

PostProcess = namedtuple("PostProcess", ["name", "function", "description"])


class _Stream:
    def __init__(self):
        self._sinks = []

    def sink(self, fn):
        self._sinks.append(fn)
        return self

    def emit(self, item):
        for fn in self._sinks:
            fn(item)


class Timer:
    def __init__(self, last, total):
        self.last, self.total = last, total

    def __repr__(self):
        return "last: %g s, total: %g s" % (self.last, self.total)


def _accepted_kwargs(kwargs, function):
    names = inspect.signature(function).parameters
    return {k: v for k, v in kwargs.items() if k in names}


class Simulation:
    """``for t, fields in Simulation(model, fields, pars, dt, tmax=...)``."""

    def __init__(self, model, fields, parameters, dt, t=0, tmax=None, id=None,
                 hook=schemes.null_hook, scheme=None, time_stepping=True, **kwargs):
        scheme = schemes.RODASPR if scheme is None else scheme
        kwargs["time_stepping"] = time_stepping
        self.id = str(uuid1())[:6] if not id else id
        self.model = model
        self.parameters = parameters
        self.fields = (fields if hasattr(fields, "uflat")
                       else model.fields_template(**fields))
        self.t = t
        self.user_dt = self.dt = dt
        self.tmax = tmax
        self.i = 0
        self.stream = _Stream()
        self._pprocesses = []
        init = scheme.__init__ if inspect.isclass(scheme) else scheme
        self._scheme = scheme(model, **_accepted_kwargs(kwargs, init))
        if time_stepping:          # always true for instances, see module doc
            self._scheme = schemes.time_stepping(
                self._scheme, **_accepted_kwargs(kwargs, schemes.time_stepping))
        self.status = "created"
        self._total_running = 0
        self._last_running = 0
        self._hook = hook
        self._iterator = self.compute()

    def _compute_one_step(self, t, fields, pars):
        fields, pars = self._hook(t, fields, pars)
        self.dt = (self.tmax - t if self.tmax and (t + self.dt >= self.tmax)
                   else self.dt)
        before = time.process_time()
        t, fields = self._scheme(t, fields, self.dt, pars, hook=self._hook)
        self._last_running = time.process_time() - before
        self._total_running += self._last_running
        return t, fields, pars

    def compute(self):
        fields, t, pars = self.fields, self.t, self.parameters
        self.stream.emit(self)
        try:
            while True:
                t, fields, pars = self._compute_one_step(t, fields, pars)
                self.i += 1
                self.t, self.fields, self.parameters = t, fields, pars
                for pprocess in self._pprocesses:
                    pprocess.function(self)
                self.stream.emit(self)
                yield self.t, self.fields
                if self.tmax and isclose(self.t, self.tmax):
                    self.status = "finished"
                    return
        except RuntimeError:
            self.status = "failed"
            raise

    def run(self, progress=False, verbose=False):
        t = fields = None
        for t, fields in self:
            pass
        if t is None:
            warnings.warn("Simulation already ended")
        return t, fields

    @property
    def post_processes(self):
        return self._pprocesses

    def add_post_process(self, name, post_process, description=""):
        self._pprocesses.append(PostProcess(name=name, function=post_process,
                                            description=description))
        self._pprocesses[-1].function(self)

    def remove_post_process(self, name):
        self._pprocesses = [p for p in self._pprocesses if p.name != name]

    @property
    def timer(self):
        return Timer(self._last_running, self._total_running)

    def __iter__(self):
        return self.compute()

    def __next__(self):
        return next(self._iterator)

    def __repr__(self):
        return "<Simulation %s t=%g dt=%g tmax=%s status=%s>" % (
            self.id, self.t, self.dt, self.tmax, self.status)

"""Host driver that iterates a scheme over output times (caller of the hot path).

Mirror of the part of the reference ``Simulation`` that the hot path is called
from (reference ``triflow/core/simulation.py:160-261``): hook before each
output step, last ``dt`` clipped to ``tmax``, ``scheme(t, fields, dt, pars,
hook=hook)``, CPU timers, post-processes, stop at ``isclose(t, tmax)``.  The
streamz / container / display plumbing of the reference (``simulation.py:
352-438``) is out of scope; ``stream`` is a minimal
emitter so user code that does ``simul.stream.sink(f)`` still runs.

Device-resident output path (``ring=K``): with a device scheme and a declarative hook the
state never leaves the GPU between outputs; every output is snapshotted into one of ``K``
pinned host buffers by an asynchronous copy on its own stream (C ABI ``tf_ring_*``) while
stepping goes on, and a consumer thread hands the finished snapshots to the ``stream``
sinks in order -- the place where the reference's container buffers and writes its netCDF
chunks (``plugins/container.py:99-137``).  Sinks then receive a :class:`Frame` (``.t``,
``.fields``, ``.i``, ``.id``, ``.parameters``) instead of the live ``Simulation``.

Quirk kept for drop-in behaviour (``simulation.py:190-197``): with the default
``time_stepping=True`` *every* scheme instance — ROS3PRw included — is wrapped
by the Richardson controller, because the reference compares an instance with
classes.  Pass ``time_stepping=False`` for fixed-step runs.
"""

import ctypes
import inspect
import threading
import time
import warnings
from collections import namedtuple
from uuid import uuid1

import numpy as np
from numpy import isclose

from . import _lib, schemes

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


Frame = namedtuple("Frame", ["id", "i", "t", "fields", "parameters"])


class OutputRing:
    """``slots`` snapshots of a device state in flight to pinned host memory
    (``tf_ring_create / push / pop / release``).  One producer, one consumer thread."""

    def __init__(self, state, slots=4):
        self.state, self.slots = state, int(slots)
        self.h = ctypes.c_void_p()
        _lib.check(_lib.lib().tf_ring_create(state.h, self.slots, ctypes.byref(self.h)))
        self.width = state.N * state.variant.lowered.nvar
        self._room = threading.Condition()
        self._in_flight = 0

    def push(self, t):
        """Snapshot the current device state; returns at once (waits only if all slots are
        in flight, i.e. the consumer is slower than the producer)."""
        with self._room:
            while self._in_flight >= self.slots:
                self._room.wait()
            self._in_flight += 1
        _lib.check(_lib.lib().tf_ring_push(self.h, float(t)))

    def pop(self, block=True):
        """Oldest finished snapshot as ``(t, array (batch, N*nvar))`` -- a view of the pinned
        slot, valid until :meth:`release` -- or ``None``."""
        data = ctypes.POINTER(ctypes.c_double)()
        t = ctypes.c_double()
        _lib.check(_lib.lib().tf_ring_pop(self.h, int(block), ctypes.byref(data), ctypes.byref(t)))
        if not data:
            return None
        arr = np.ctypeslib.as_array(data, shape=(self.state.batch, self.width))
        return t.value, arr

    def release(self):
        _lib.check(_lib.lib().tf_ring_release(self.h))
        with self._room:
            self._in_flight -= 1
            self._room.notify()

    def close(self):
        if self.h:
            _lib.lib().tf_ring_destroy(self.h)
            self.h = None


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
                 hook=schemes.null_hook, scheme=None, time_stepping=True, ring=0, **kwargs):
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
        # device-resident output path: needs a device scheme called directly (no host-side
        # Richardson wrapper around it) and a hook that runs on the device
        self._ring_slots = int(ring) if (ring and hasattr(self._scheme, "_bind")
                                         and self._scheme._on_device(hook)) else 0
        self._ring = self._consumer = None
        self.frames_emitted = 0
        if self._ring_slots:
            self._scheme.lazy = True
        self._iterator = self.compute()

    # ---- device-resident output path
    def _consume(self, template):
        """Consumer thread: finished snapshots -> stream sinks, in order."""
        while True:
            item = self._pending.get()
            if item is None:
                return
            i, pars = item
            t, data = self._ring.pop(block=True)
            if hasattr(template, "with_uflat"):
                fields = template.with_uflat(data[0])        # one pass over the snapshot
            else:
                fields = template.copy()
                fields.fill(data[0])
            self._ring.release()
            self.stream.emit(Frame(self.id, i, t, fields, pars))
            self.frames_emitted += 1

    def _ring_push(self, t, fields, pars):
        if self._ring is None:
            import queue
            self._ring = OutputRing(self._scheme._state, self._ring_slots)
            self._pending = queue.Queue()
            template = fields._template if hasattr(fields, "_template") else fields
            self._consumer = threading.Thread(target=self._consume, args=(template,), daemon=True)
            self._consumer.start()
        self._ring.push(t)
        self._pending.put((self.i, pars))

    def drain(self):
        """Wait until every pushed output has reached the sinks (end of a ring run)."""
        if self._consumer is not None:
            self._pending.put(None)
            self._consumer.join()
            self._consumer = None
            self._ring.close()
            self._ring = None

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
                if self._ring_slots and getattr(fields, "is_resident", None) \
                        and fields.is_resident(self._scheme):
                    self._ring_push(t, fields, pars)       # asynchronous: copy + sinks overlap stepping
                else:
                    self.stream.emit(self)
                yield self.t, self.fields
                if self.tmax and isclose(self.t, self.tmax):
                    self.drain()
                    self.status = "finished"
                    return
        except RuntimeError:
            self.drain()
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

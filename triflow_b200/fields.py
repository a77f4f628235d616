"""Host-side field container (the layout contract of the hot path).

Mirrors the duck-typed surface of the reference's ``BaseFields``
(reference ``triflow/core/fields.py:21-189``) that the F/J routines, the
schemes and ``Simulation`` rely on — ``fields[name].values``, item
assignment, ``uflat``, ``fill``, ``copy``, ``dependent_variables``,
``helper_functions``, ``size``, ``keys`` — without xarray (not installed on
the GPU boxes).  Only the 1-D case of the reference is covered, which is the
only one its ``Model`` can produce (``fields.py:79-105``).

Layout contract (reference ``fields.py:146-159,173-183``)::

    uflat[i * nvar + e] == fields[dependent_variables[e]][i]

``fill`` is the exact inverse.
"""

import ctypes

import numpy as np


class FieldArray(np.ndarray):
    """ndarray that also answers ``.values`` like an ``xarray.DataArray``.

    The reference reads ``fields[key].values`` (``core/routines.py:38-43``)
    and hooks write ``fields["U"][0] = 1`` (``README.md:126-129``); both work
    on this view without copying.
    """

    @property
    def values(self):
        return self.view(np.ndarray)


def _as_field(a):
    return np.array(a, dtype=np.float64, copy=True).view(FieldArray)


class BaseFields:
    """Per-model container; specialised by :meth:`factory1D`."""

    dependent_variables = []
    helper_functions = []
    _coords = ("x",)

    @staticmethod
    def factory1D(dependent_variables, helper_functions):
        """Build the container class of a model (``fields.py:79-105``)."""
        deps = [str(n) for n in dependent_variables]
        helps = [str(n) for n in helper_functions]
        cls = type("Field", (BaseFields,), {})
        cls.dependent_variables = deps
        cls.helper_functions = helps
        cls.dependent_variables_info = [(n, ("x",)) for n in deps]
        cls.helper_functions_info = [(n, ("x",)) for n in helps]
        cls._keys = tuple(deps + helps)
        return cls

    def __init__(self, **inputs):
        # KeyError on a missing variable, like the reference (fields.py:107-112)
        self._data = {"x": _as_field(inputs["x"])}
        for key in self._keys:
            arr = _as_field(inputs[key])
            if arr.shape != self._data["x"].shape:
                arr = (arr + np.zeros_like(self._data["x"])).view(FieldArray)
            self._data[key] = arr

    # -- mapping surface ---------------------------------------------------
    def keys(self):
        return [*self._coords, *self._keys]

    def __getitem__(self, key):
        return self._data[key]

    def __setitem__(self, key, value):
        if key not in self._data:
            raise KeyError(key)
        self._data[key][...] = value

    def __getattr__(self, name):
        data = self.__dict__.get("_data")
        if data is not None and name in data:
            return data[name]
        raise AttributeError(name)

    def __contains__(self, key):
        return key in self._data

    def __iter__(self):
        return iter(self.keys())

    # -- reference API -----------------------------------------------------
    @property
    def size(self):
        return self._data["x"].size

    @property
    def uflat(self):
        """Flat **copy** of the dependent variables, node-major/variable-minor."""
        return np.stack([self._data[k].view(np.ndarray)
                         for k in self.dependent_variables], axis=1).reshape(-1)

    def fill(self, uflat):
        """Inverse of :attr:`uflat` (``fields.py:173-183``)."""
        rarray = np.asarray(uflat, dtype=np.float64).reshape((self.size, -1))
        for e, var in enumerate(self.dependent_variables):
            self._data[var][...] = rarray[:, e]

    def copy(self, deep=True):
        new = object.__new__(type(self))
        new._data = {k: (v.copy() if deep else v) for k, v in self._data.items()}
        return new

    __copy__ = copy

    def with_uflat(self, uflat):
        """New container that SHARES ``x`` and the helper fields with this one and whose
        unknowns are strided views of one private copy of ``uflat`` (one pass over the data
        instead of ``copy()`` + ``fill()``; used by the output ring's consumer)."""
        src = np.ascontiguousarray(uflat, dtype=np.float64)
        flat = np.empty(src.size, dtype=np.float64)
        # (a foreign call: the GIL is released while the snapshot is copied, so the thread
        #  that keeps stepping is not held up by the consumer)
        ctypes.memmove(flat.ctypes.data, src.ctypes.data, src.nbytes)
        flat = flat.reshape(self.size, -1)
        new = object.__new__(type(self))
        new._data = dict(self._data)
        for e, var in enumerate(self.dependent_variables):
            new._data[var] = flat[:, e].view(FieldArray)
        return new

    def to_df(self):
        import pandas as pd
        return pd.DataFrame({k: self._data[k].view(np.ndarray) for k in self._keys},
                            index=self._data["x"].view(np.ndarray))

    def to_csv(self, path):
        self.to_df().to_csv(path)

    def __reduce__(self):
        return (_rebuild, (self.dependent_variables, self.helper_functions,
                           {k: v.view(np.ndarray) for k, v in self._data.items()}))

    def __repr__(self):
        return "<Fields N=%d vars=%s helpers=%s>" % (
            self.size, self.dependent_variables, self.helper_functions)


def _rebuild(deps, helps, data):
    return BaseFields.factory1D(deps, helps)(**data)


class LazyFields:
    """Result of a device-resident scheme call (opt-in, ``scheme(model, lazy=True)``).

    Behaves like the fields container but downloads the unknowns only when they are
    first touched.  Passing it back to the scheme that produced it *consumes* it: the
    device state advances in place and nothing crosses PCIe.  Touching a consumed
    object raises ``RuntimeError`` (read the values before stepping again, or keep
    ``lazy=False``, which mirrors the reference exactly)."""

    def __init__(self, template, owner, generation):
        object.__setattr__(self, "_template", template)      # x / helpers / layout
        object.__setattr__(self, "_owner", owner)
        object.__setattr__(self, "_generation", generation)
        object.__setattr__(self, "_real", None)

    def _load(self):
        if self._real is None:
            owner = self._owner
            if owner._generation != self._generation:
                raise RuntimeError(
                    "these lazy fields were consumed by a later device step; read them "
                    "before stepping again or construct the scheme with lazy=False")
            real = self._template.copy()
            real.fill(owner._state.download()[0])
            object.__setattr__(self, "_real", real)
        return self._real

    def is_resident(self, owner):
        """True when the device state of ``owner`` still holds exactly these values."""
        return (self._real is None and owner is self._owner
                and owner._generation == self._generation)

    # everything else is the container API, after materialisation
    def __getitem__(self, key):
        if self._real is None and (key == "x" or key in self._template.helper_functions):
            return self._template[key]           # no download needed for grid / helper fields
        return self._load()[key]

    def __setitem__(self, key, value):
        self._load()[key] = value

    def __getattr__(self, name):
        if name in ("dependent_variables", "helper_functions", "size", "keys"):
            return getattr(self._template, name)
        return getattr(self._load(), name)

    def __iter__(self):
        return iter(self._template.keys())

    def __contains__(self, key):
        return key in self._template

    def copy(self, deep=True):
        return self._load().copy(deep)

    def __repr__(self):
        return "<LazyFields %s>" % ("on device" if self._real is None else repr(self._real))

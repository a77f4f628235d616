"""The CUDA compiler plugin: ``Model(..., compiler=cuda_compiler)``.

Drop-in for the reference's ``numpy_compiler`` / ``theano_compiler``
(reference ``triflow/core/compilers.py:11-224``): called as
``F_function, J_function = compiler(model)`` (``core/model.py:299-300``) and the
two callables obey the ufunc contract of ``core/routines.py:37-45,82-91``:

    F_function(x, *dep_vars, *helpers, *pars, periodic) -> ndarray (N*nvar,)
    J_function(x, *dep_vars, *helpers, *pars, periodic) -> scipy.sparse.csc_matrix

Both evaluate on the GPU (kernels ``tf_k_eval_F`` / ``tf_k_eval_J`` of the
model's cubin); only the placement of the J values into CSC — the reference's
index rule ``compilers.py:303-331`` — is host numpy, because the compatibility
path has to hand back a SciPy matrix.  The schemes of :mod:`triflow_b200.schemes`
do not go through these callables: they keep the state on the device and use the
fused factor / solve kernels; they find the compiled model through
``model._cuda``.
"""

import ctypes as C
import os

import numpy as np
from scipy.sparse import csc_matrix

from . import _lib, codegen


def default_chunk_nodes(nvar, half_width):
    if os.environ.get("TF_CHUNK_NODES"):            # tuning knob
        return int(os.environ["TF_CHUNK_NODES"])
    beta = half_width * nvar + nvar - 1
    m = 8
    while m * nvar > 8 and m > 1 and (m // 2) * nvar >= beta:
        m //= 2
    return m


def value_kind(shape, batch, N, name="value", field=False):
    """How a parameter / field value is laid out for ``batch`` systems of ``N`` nodes:
    ``"scalar"``, ``"member"`` (one value per system), ``"node"`` (per node, shared) or
    ``"member_node"``.  A 1-D PARAMETER array is ambiguous when ``batch == N``: say which with
    a ``(batch, 1)`` or ``(1, N)`` shape (a 1-D ``field`` is a function on the grid: per node)."""
    shape = tuple(shape)
    if len(shape) == 0:
        return "scalar"
    if len(shape) == 1:
        if field and shape[0] == N:
            return "node"
        if batch > 1 and batch == N and shape[0] == N:
            raise ValueError("%s: a 1-D array of length %d is ambiguous when batch == N; pass "
                             "shape (batch, 1) for one value per member or (1, N) per node" % (name, N))
        if batch > 1 and shape[0] == batch:
            return "member"
        if shape[0] == N:
            return "node"
    elif len(shape) == 2 and shape[0] in (1, batch):
        if shape[1] == 1:
            return "member" if shape[0] == batch and batch > 1 else "scalar"
        if shape[1] == N:
            return "member_node" if shape[0] == batch and batch > 1 else "node"
    raise ValueError("%s: shape %s fits neither the batch (%d) nor the grid (%d)" % (name, shape, batch, N))


class DeviceState:
    """Fields of ``batch`` systems resident on the device."""

    def __init__(self, cmodel, variant, N, batch, periodic, ctx=None):
        self.cmodel, self.variant = cmodel, variant
        self.ctx = cmodel.ctx if ctx is None else ctx     # stream the state's work runs on
        self.N, self.batch, self.periodic = int(N), int(batch), bool(periodic)
        self.h = C.c_void_p()
        handle = variant.handle                            # loads the cubin on first use
        _lib.check(_lib.lib().tf_state_create(self.ctx, handle, self.N, self.batch,
                                              int(self.periodic), C.byref(self.h)))
        self._x = None
        self._consts = None

    def close(self):
        if self.h:
            _lib.lib().tf_state_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def upload(self, x=None, u=None, helpers=None, nodepars=None, consts=None):
        def prep(a):
            return None if a is None else np.ascontiguousarray(a, dtype=np.float64)
        x, u, helpers, nodepars, consts = map(prep, (x, u, helpers, nodepars, consts))
        _lib.check(_lib.lib().tf_state_upload(self.h, _lib.dptr(x), _lib.dptr(u),
                                              _lib.dptr(helpers), _lib.dptr(nodepars),
                                              _lib.dptr(consts)))

    def set_inputs(self, x, fields_by_name, pars):
        """Upload grid, helper fields and parameters (not the unknowns)."""
        L = self.variant.lowered
        x = np.asarray(x, dtype=np.float64)
        upl = {}
        if self._x is None or not np.array_equal(self._x, x):
            upl["x"] = x
            self._x = x.copy()
        dx = (x[-1] - x[0]) / (x.size - 1)                       # compilers.py:234-237
        table = L.uniform_table(dx, pars, self.batch)
        if self._consts is None or not np.array_equal(self._consts, table):
            upl["consts"] = table
            self._consts = table
        if L.nhelp:
            upl["helpers"] = np.stack(
                [np.broadcast_to(np.asarray(fields_by_name[h], float), (self.batch, self.N))
                 for h in L.fields[L.nvar:]], axis=1)
        if L.node_pars:
            upl["nodepars"] = np.stack(
                [np.broadcast_to(np.asarray(pars[p], float), (self.batch, self.N))
                 for p in L.node_pars], axis=1)
        if upl:
            self.upload(**upl)

    def download(self, out=None):
        nv = self.variant.lowered.nvar
        if out is None:
            out = np.empty((self.batch, self.N * nv))
        _lib.check(_lib.lib().tf_state_download(self.h, _lib.dptr(out)))
        return out

    def eval_F(self):
        out = np.empty((self.batch, self.N * self.variant.lowered.nvar))
        _lib.check(_lib.lib().tf_eval_F(self.h, _lib.dptr(out)))
        return out

    def eval_J(self):
        out = np.empty((self.batch, self.N, self.variant.lowered.nnz))
        _lib.check(_lib.lib().tf_eval_J(self.h, _lib.dptr(out)))
        return out

    def status(self):
        out = (C.c_int * self.batch)()
        _lib.check(_lib.lib().tf_state_status(self.h, out))
        return np.array(out[:])


class Variant:
    """One lowering of the model (a given set of per-node parameters) + its cubin.
    Lowering and the nvcc build need no GPU; the cubin is loaded on first use."""

    def __init__(self, cmodel, node_pars):
        self.cmodel = cmodel
        self.lowered = L = codegen.lower(cmodel.model, node_pars)
        self.chunk_nodes = cmodel.chunk_nodes or default_chunk_nodes(L.nvar, L.half_width)
        self.warps = cmodel.warps
        self.cubin_path = _lib.build_cubin(L.header, self.chunk_nodes, self.warps,
                                           cmodel.fast_div)
        self._handle = None

    @property
    def handle(self):
        if self._handle is None:
            L = self.lowered
            with open(self.cubin_path, "rb") as f:
                self._image = f.read()
            desc = _lib.ModelDesc(L.nvar, L.nhelp, L.half_width, L.nnz, L.n_const,
                                  len(L.node_pars), int(L.uses_x), self.chunk_nodes,
                                  self.warps)
            h = C.c_void_p()
            _lib.check(_lib.lib().tf_model_load(self.cmodel.ctx, self._image,
                                                len(self._image), C.byref(desc), C.byref(h)))
            self._handle = h
        return self._handle


class CompiledModel:
    """Everything the GPU path knows about one ``Model``."""

    def __init__(self, model, fast_div=True, chunk_nodes=None, warps=8, device=None):
        self.model = model
        self.fast_div = bool(fast_div)
        self.chunk_nodes = chunk_nodes
        self.warps = warps
        self.device = device
        self._ctx = None
        self._variants = {}
        self._states = {}

    @property
    def ctx(self):
        if self._ctx is None:
            self._ctx = _lib.context(self.device)
        return self._ctx

    def node_pars_of(self, pars, N, batch=1):
        """Parameters given as per-node arrays (reference core/routines.py:40 accepts
        scalars or (N,) arrays).  With batch > 1 a (batch,) array is one value per
        system (uniform inside a system); (N,) / (batch, N) arrays are per node."""
        out = []
        for p in self.model._pars:
            if value_kind(np.shape(pars[p]), batch, N, "parameter %r" % p) in ("node", "member_node"):
                out.append(p)
        return tuple(out)

    def variant(self, node_pars=()):
        node_pars = tuple(node_pars)
        if node_pars not in self._variants:
            self._variants[node_pars] = Variant(self, node_pars)
        return self._variants[node_pars]

    def new_state(self, pars, N, batch, periodic, ctx=None):
        v = self.variant(self.node_pars_of(pars, N, batch))
        return DeviceState(self, v, N, batch, periodic, ctx=ctx)

    def cached_state(self, pars, N, periodic):
        v = self.variant(self.node_pars_of(pars, N, 1))
        key = (id(v), int(N), bool(periodic))
        if key not in self._states:
            self._states[key] = DeviceState(self, v, N, 1, periodic)
        return self._states[key]

    # ---- reference ufunc contract ---------------------------------------
    def _split_args(self, args):
        m = self.model
        names = [*m._indep_vars, *m._dep_vars, *m._help_funcs, *m._pars, "periodic"]
        named = dict(zip(names, args))
        pars = {p: named[p] for p in m._pars}
        return named, pars, bool(named["periodic"])

    def _prepare(self, args):
        named, pars, periodic = self._split_args(args)
        x = np.asarray(named["x"], dtype=np.float64)
        st = self.cached_state(pars, x.size, periodic)
        st.set_inputs(x, named, pars)
        nv = self.model._nvar
        u = np.stack([np.asarray(named[v], dtype=np.float64) for v in self.model._dep_vars],
                     axis=1).reshape(1, x.size * nv)
        st.upload(u=u)
        return st, x.size, periodic

    def F_function(self, *args):
        st, N, _ = self._prepare(args)
        return st.eval_F()[0]

    def J_function(self, *args):
        st, N, periodic = self._prepare(args)
        vals = st.eval_J()[0]                                  # (N, nnz)
        L = st.variant.lowered
        nv = L.nvar
        i = np.arange(N)[:, None]
        eq, var, off = (np.asarray(a)[None, :] for a in (L.j_eq, L.j_var, L.j_off))
        j = i + off
        j = j % N if periodic else np.clip(j, 0, N - 1)
        return csc_matrix((vals.reshape(-1), ((i * nv + eq).reshape(-1),
                                              (j * nv + var).reshape(-1))),
                          shape=(N * nv, N * nv))


def cuda_compiler(model, **options):
    """Compiler plugin entry point (reference ``core/model.py:299-300``)."""
    cm = CompiledModel(model, **options)
    model._cuda = cm
    return cm.F_function, cm.J_function


def make_cuda_compiler(**options):
    """``Model(..., compiler=make_cuda_compiler(fast_div=True))``."""
    def compiler(model):
        return cuda_compiler(model, **options)
    return compiler

"""Multi-GPU: one process per GPU.

* Ensembles are sharded by member, no data-path collective.
* ONE grid larger than a GPU's resident tiles is cut into slabs, one per GPU
  (:class:`SlabGrid`): the grid-resident step kernel runs on every GPU at once and reads its
  neighbours' halo values, scan records and border-block words straight from the peer GPU's
  memory over NVLink -- the exchange is fused into the step kernel, no collective call.

``torch.distributed`` is used for the rendezvous (incl. the exchange of the 64-byte IPC
handles), the barrier around timed regions and the final gather only."""

import os

import numpy as np


def world():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def shard(n_members, rank, world_size):
    """Contiguous block of members ``[lo, hi)`` owned by ``rank``."""
    base, rem = divmod(int(n_members), int(world_size))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def init(backend=None):
    """Join the process group described by the torchrun environment."""
    import torch
    import torch.distributed as dist
    rank, ws = world()
    if ws == 1 or dist.is_initialized():
        return rank, ws
    if backend is None:
        backend = "nccl" if torch.cuda.is_available() else "gloo"
    kw = {}
    if backend == "nccl":
        local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(local)
        kw["device_id"] = torch.device("cuda", local)
    dist.init_process_group(backend=backend, rank=rank, world_size=ws, **kw)
    return rank, ws


def finalize():
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        dist.destroy_process_group()


def barrier():
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        dist.barrier()


def max_over_ranks(value):
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return float(value)
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    t = torch.tensor([float(value)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value):
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return float(value)
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    t = torch.tensor([float(value)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def gather_members(local, n_members, out=None):
    """Final gather of per-member rows ``local`` (n_local, width) to rank 0 in member order.
    ``out``: optional ``(n_members, width)`` numpy array on rank 0 for the result (page-locked,
    ``_lib.pinned_empty``, for a full-speed copy).
    ``local``: a numpy array, or -- NCCL backend -- a ``torch`` CUDA tensor
    (``Ensemble.download_to_torch``): then nothing crosses PCIe except rank 0's single copy of
    the result to the host.  Returns the full array on rank 0, ``None`` elsewhere."""
    import torch
    import torch.distributed as dist
    out_host = out
    on_device = isinstance(local, torch.Tensor)
    if not (dist.is_available() and dist.is_initialized()):
        return local.cpu().numpy() if on_device else np.asarray(local)
    rank, ws = dist.get_rank(), dist.get_world_size()
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    width = local.shape[1]
    sizes = [shard(n_members, r, ws) for r in range(ws)]
    nmax = max(hi - lo for lo, hi in sizes)
    if on_device and local.shape[0] == nmax and dev == "cuda":
        buf = local                                  # equal shards: send the state buffer itself
    else:
        buf = torch.zeros((nmax, width), dtype=torch.float64, device=dev)
        src = local if on_device else torch.from_numpy(np.ascontiguousarray(local))
        buf[: local.shape[0]] = src.to(dev)
    out = [torch.empty_like(buf) for _ in range(ws)] if rank == 0 else None
    dist.gather(buf, out, dst=0)
    if rank != 0:
        return None
    if dev == "cuda":                                # one device -> host copy of the result
        res = np.empty((n_members, width)) if out_host is None else out_host
        full = torch.from_numpy(res)
        for o, (lo, hi) in zip(out, sizes):
            full[lo:hi].copy_(o[: hi - lo])
        torch.cuda.synchronize()
        return res
    return np.concatenate([o[: hi - lo].cpu().numpy() for o, (lo, hi) in zip(out, sizes)])


def slab_partition(N, nranks, tile_nodes, tiles_local):
    """``[(node_off, n_local)]`` of every rank for a grid of ``N`` nodes cut into slabs of
    ``tiles_local`` tiles of ``tile_nodes`` nodes (mirrors ``create_state`` of tf_host.cu)."""
    out = []
    for r in range(nranks):
        off = r * tiles_local * tile_nodes
        out.append((off, max(0, min(N - off, tiles_local * tile_nodes))))
    return out


def gather_slabs(local, partition):
    """Final gather of a slab grid: rank r holds ``partition[r][1]`` consecutive nodes starting
    at ``partition[r][0]``.  Returns the whole grid on rank 0, ``None`` elsewhere."""
    import torch
    import torch.distributed as dist
    local = np.ascontiguousarray(local, dtype=np.float64)
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return local
    rank, ws = dist.get_rank(), dist.get_world_size()
    if local.size != partition[rank][1]:
        raise ValueError("rank %d holds %d nodes, its slab has %d" % (rank, local.size, partition[rank][1]))
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    nmax = max(n for _, n in partition)
    buf = torch.zeros(nmax, dtype=torch.float64, device=dev)
    buf[:local.size] = torch.from_numpy(local).to(dev)
    out = [torch.empty_like(buf) for _ in range(ws)] if rank == 0 else None
    dist.gather(buf, out, dst=0)
    if rank != 0:
        return None
    full = np.concatenate([o[:n].cpu().numpy() for o, (_, n) in zip(out, partition)])
    offs = [off for off, _ in partition]
    assert offs == sorted(offs) and offs[0] == 0
    return full


class _SlabState:
    """This rank's slab of a grid spread over several GPUs (``tf_state_create_slab``)."""

    def __init__(self, cmodel, N, periodic, rank, nranks):
        import ctypes as C
        from . import _lib
        self.cmodel, self.variant = cmodel, cmodel.variant(())
        self.ctx = cmodel.ctx
        self.N, self.rank, self.nranks = int(N), int(rank), int(nranks)
        self.h = C.c_void_p()
        L = _lib.lib()
        _lib.check(L.tf_state_create_slab(self.ctx, self.variant.handle, self.N, int(bool(periodic)),
                                          self.rank, self.nranks, C.byref(self.h)))
        a, b, c, d = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        _lib.check(L.tf_state_slab_info(self.h, C.byref(a), C.byref(b), C.byref(c), C.byref(d)))
        self.node_off, self.n_local, self.tiles_local, self.tiles_total = a.value, b.value, c.value, d.value

    def export(self):
        import ctypes as C
        from . import _lib
        buf = C.create_string_buffer(64)
        _lib.check(_lib.lib().tf_state_slab_export(self.h, buf))
        return buf.raw

    def attach(self, handles):
        from . import _lib
        _lib.check(_lib.lib().tf_state_slab_attach(self.h, b"".join(handles)))

    def upload(self, u=None, consts=None):
        from . import _lib
        prep = lambda a: None if a is None else np.ascontiguousarray(a, dtype=np.float64)  # noqa: E731
        u, consts = prep(u), prep(consts)
        _lib.check(_lib.lib().tf_state_upload(self.h, None, _lib.dptr(u), None, None, _lib.dptr(consts)))

    def download(self):
        from . import _lib
        out = np.empty(self.n_local)
        _lib.check(_lib.lib().tf_state_download(self.h, _lib.dptr(out)))
        return out

    def status(self):
        import ctypes as C
        from . import _lib
        out = (C.c_int * 1)()
        _lib.check(_lib.lib().tf_state_status(self.h, out))
        return int(out[0])

    def close(self):
        from . import _lib
        if self.h:
            _lib.lib().tf_state_destroy(self.h)
            self.h = None


class SlabGrid:
    """One grid of ``N`` nodes stepped by all GPUs of the process group together.

    Every rank constructs it with the same arguments (``x`` and ``fields`` of the WHOLE grid);
    rank ``r`` keeps nodes ``node_off .. node_off + n_local - 1`` on its GPU.  ``step`` launches the same cooperative step kernel on every GPU;
    inside a step the GPUs are coupled only through tagged words read from the neighbour's
    memory over NVLink.  Fixed steps, scalar model, uniform parameters, tableaux of <= 3 stages;
    ``hook``: ``schemes.Dirichlet`` on non-periodic grids.

    ``devices``: single-process form -- one Python process drives the listed GPUs (the object
    then holds the slabs of all ranks and ``upload`` / ``download`` take the whole grid).
    """

    def __init__(self, model, scheme, x, fields, pars, devices=None, hook=None):
        from . import _lib
        from .compiler import CompiledModel
        from .schemes import Dirichlet
        self.model, self.scheme = model, scheme
        x = np.asarray(x, dtype=np.float64)
        self.N = x.size
        periodic = bool(pars["periodic"])
        (var,) = model._dep_vars
        u = np.asarray(fields[var], dtype=np.float64)
        self.t = 0.0
        self._deferred = None
        self._local = devices is not None
        dx = (x[-1] - x[0]) / (x.size - 1)                       # compilers.py:234-237
        if self._local:
            self._cms = []
            for d in devices:                 # own context (own stream, asynchronous) per GPU
                cm = CompiledModel(model, device=d)
                cm._ctx = _lib.new_context(d)
                self._cms.append(cm)
            n = len(devices)
            self.states = [_SlabState(cm, self.N, periodic, r, n) for r, cm in enumerate(self._cms)]
            import ctypes as C
            arr = (C.c_void_p * n)(*[s.h for s in self.states])
            for s in self.states:
                _lib.check(_lib.lib().tf_state_slab_attach_local(s.h, arr))
                _lib.check(_lib.lib().tf_ctx_set_async(s.ctx, 1))
            self.rank, self.nranks = 0, n
        else:
            import torch.distributed as dist
            self.rank, self.nranks = (dist.get_rank(), dist.get_world_size()) if dist.is_initialized() else (0, 1)
            # every step below is collective: a failure on one rank must fail all of them, or the
            # others would wait for it in the next collective / spin on its words
            err, st = None, None
            try:
                st = _SlabState(model._cuda, self.N, periodic, self.rank, self.nranks)
                handle = st.export()
            except Exception as e:  # noqa: BLE001
                err, handle = e, b""
            self.states = [st] if st is not None else []
            handles = [None] * self.nranks
            if self.nranks > 1:
                dist.all_gather_object(handles, handle)
            else:
                handles = [handle]
            if err is None and all(len(h) == 64 for h in handles):
                try:
                    st.attach(handles)
                except Exception as e:  # noqa: BLE001
                    err = e
            elif err is None:
                err = RuntimeError("slab grid: another rank could not create its slab")
            if max_over_ranks(0.0 if err is None else 1.0) > 0:
                self.close()
                raise err if err is not None else RuntimeError("slab grid: another rank failed to attach")
        for s in self.states:
            table = s.variant.lowered.uniform_table(dx, pars, 1)
            s.upload(consts=table)
            if isinstance(hook, Dirichlet):            # declarative hook, non-periodic grids
                left, right = hook.values[var]
                _lib.check(_lib.lib().tf_hook_set_dirichlet(
                    s.h, 0, left is not None, 0.0 if left is None else float(left),
                    right is not None, 0.0 if right is None else float(right)))
            elif hook is not None and getattr(hook, "__name__", "") != "null_hook":
                raise ValueError("slab grids take declarative hooks (schemes.Dirichlet) only")
        self.upload(u)

    # -- this rank's part of the grid
    @property
    def node_off(self):
        return self.states[0].node_off

    @property
    def n_local(self):
        return self.states[0].n_local

    def upload(self, u):
        """``u``: the whole grid (every rank takes its slice) or, in the one-process-per-GPU
        form, just this rank's ``n_local`` nodes."""
        u = np.asarray(u, dtype=np.float64).ravel()
        for s in self.states:
            part = u if (u.size == s.n_local and not self._local and u.size != self.N) \
                else u[s.node_off:s.node_off + s.n_local]
            s.upload(u=part)
        self.sync()
        barrier()          # every rank's edge words are in place before anybody steps

    def step(self, dt, n_steps=1):
        from . import _lib
        L = _lib.lib()
        for s in self.states:
            try:
                _lib.check(L.tf_scheme_step(s.h, self.scheme.handle, float(dt), int(n_steps), None))
            except Exception as e:  # noqa: BLE001  (reported by the next sync(), on every rank)
                self._deferred = self._deferred or ("rank %d: %s" % (s.rank, e))
        self.t += n_steps * dt

    def sync(self):
        """Wait for the steps issued so far.  Collective in the one-process-per-GPU form: every
        rank learns whether ANY rank failed (a time-out on one rank is every rank's error)."""
        from . import _lib
        bad, self._deferred = self._deferred, None
        for s in self.states:
            try:
                _lib.check(_lib.lib().tf_ctx_sync(s.ctx))
                st = s.status()
            except Exception as e:  # noqa: BLE001
                bad = bad or ("rank %d: %s" % (s.rank, e))
                continue
            if st and bad is None:
                bad = ("rank %d reports status %d (bit 2: a tile timed out waiting for a "
                       "neighbour; bits 0/1/3: factorisation)" % (s.rank, st))
        anybad = bad is not None
        if not self._local and self.nranks > 1:
            anybad = max_over_ranks(1.0 if anybad else 0.0) > 0
        if anybad:
            raise RuntimeError("slab grid: " + (bad or "another rank failed"))

    def download(self):
        """This rank's nodes (single-process form: the whole grid)."""
        self.sync()
        parts = [s.download() for s in self.states]
        self.sync()
        return np.concatenate(parts) if self._local else parts[0]

    def gather(self):
        """The whole grid on rank 0 (``None`` elsewhere)."""
        local = self.download()
        if self._local or self.nranks == 1:
            return local
        return gather_slabs(local, self.partition())

    def partition(self):
        """``[(node_off, n_local)]`` of every rank."""
        if self._local:
            return [(t.node_off, t.n_local) for t in self.states]
        s = self.states[0]
        if self.nranks == 1:
            return [(0, self.N)]
        tile_nodes = s.node_off // (s.rank * s.tiles_local) if s.rank else s.n_local // s.tiles_local
        return slab_partition(self.N, self.nranks, tile_nodes, s.tiles_local)

    def close(self):
        """Collective in the one-process-per-GPU form: nobody frees a record area that a peer's
        kernel may still be reading."""
        if not self._local and self.nranks > 1:
            try:
                for s in self.states:
                    from . import _lib
                    _lib.lib().tf_ctx_sync(s.ctx)
                barrier()
            except Exception:  # noqa: BLE001
                pass
        for s in self.states:
            s.close()
        self.states = []

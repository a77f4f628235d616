"""One grid over several GPUs (triflow_b200.distributed.SlabGrid, tf_state_create_slab).

The reference has no analogue (one host, sparse storage: source_doc/source/user_guide.rst:183-187),
so the checker is the same grid stepped on ONE GPU by the per-kernel pipeline (which is pinned
to the reference's golden trajectories by tests/test_gpu_parity.py) and the CPU oracle.

* 1 GPU (what the driver's box has): the several-GPU kernel tf_k_gridstep_mr with one rank --
  slab bookkeeping, dead tiles, seed words, the C ABI -- against the pipeline and the oracle.
* >= 2 GPUs: two slabs, (a) one process driving both GPUs (peer pointers), (b) one process per
  GPU under torchrun (IPC-mapped record areas), both against the single-GPU result.
"""
import os
import subprocess
import sys

import numpy as np
import pytest

from helpers import rel_traj_err

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FX = dict(time_stepping=False)


def _ngpus():
    import ctypes
    try:
        rt = ctypes.CDLL("libcuda.so.1")
        n = ctypes.c_int()
        if rt.cuInit(0) != 0 or rt.cuDeviceGetCount(ctypes.byref(n)) != 0:
            return 0
        return n.value
    except OSError:
        return 0


def _case(name, N):
    from triflow_b200 import schemes as S, workloads as W
    from test_gpu_parity import gmodel
    if name == "ks":
        c = W.kuramoto(N)
        m = gmodel("ks")
        return m, S.ROS3PRw(m, **FX), c
    if name == "ks_edge":
        c = W.kuramoto(N)
        c["pars"] = dict(periodic=False)
        m = gmodel("ks")
        return m, S.ROS3PRw(m, **FX), c
    if name == "ks_theta":
        c = W.kuramoto(N)
        m = gmodel("ks")
        return m, S.Theta(m, theta=0.5), c
    if name == "heat":                        # periodic, the border fill reaches every tile
        x = np.linspace(0, 10, N)
        m = gmodel("heat")
        return m, S.ROS3PRw(m, **FX), dict(x=x, fields=dict(T=np.cos(2 * np.pi * x / 10)),
                                           pars=dict(k=1.0, periodic=True), dt=0.5)
    if name == "burgers":
        m = gmodel("burgers_up1")
        return m, S.ROS2(m), W.burgers(N, 1)
    if name == "advdiff":                     # non-periodic, Dirichlet hook at both ends
        m = gmodel("advdiff")
        c = W.readme(N)
        c["dt"] = 0.01
        c["hook"] = S.Dirichlet(U=(1.0, 0.0))
        return m, S.ROS3PRw(m, **FX), c
    raise KeyError(name)


def _single(m, sch, c, steps):
    from triflow_b200.ensemble import Ensemble
    from triflow_b200 import schemes as S
    ens = Ensemble(m, sch, c["x"], c["fields"], c["pars"], hook=c.get("hook", S.null_hook), batch=1)
    ens.set_fusion(False)                      # the per-kernel pipeline
    ens.step(c["dt"], steps)
    u = ens.download()[0].copy()
    ens.state.close()
    return u


def _slab(m, sch, c, steps, devices):
    from triflow_b200.distributed import SlabGrid
    g = SlabGrid(m, sch, c["x"], c["fields"], c["pars"], devices=devices, hook=c.get("hook"))
    g.step(c["dt"], 1)
    g.step(c["dt"], steps - 1)
    u = g.gather()
    part = g.partition()
    info = (g.states[0].tiles_local, g.states[0].tiles_total)
    g.close()
    return u, part, info


CASES = [("ks", 20000), ("ks", 50001), ("ks", 2049), ("ks_edge", 30000), ("ks_theta", 70001),
         ("heat", 20000), ("burgers", 40000), ("advdiff", 20000)]


@pytest.mark.parametrize("name,N", CASES)
def test_slab_kernel_one_rank_equals_pipeline(name, N):
    """The several-GPU step kernel with a single slab: same answer as the pipeline."""
    m, sch, c = _case(name, N)
    u, part, (tl, tt) = _slab(m, sch, c, 5, [0])
    assert part == [(0, N)] and tl == tt
    assert np.isfinite(u).all()
    assert rel_traj_err(u, _single(m, sch, c, 5)) <= 5e-10


def test_slab_kernel_one_rank_vs_oracle():
    from oracle import schemes as O
    from test_gpu_parity import omodel
    m, sch, c = _case("ks", 4096)
    u, _, _ = _slab(m, sch, c, 10, [0])
    om = omodel("ks")
    f = om.fields_template(x=c["x"], **c["fields"])
    osch, t = O.ROS3PRw(om, **FX), 0.0
    for _ in range(10):
        t, f = osch(t, f, c["dt"], c["pars"])
    assert rel_traj_err(u, f.uflat) <= 1e-8


def test_slab_state_rejects_what_it_cannot_do():
    import ctypes as C
    from triflow_b200 import _lib
    from triflow_b200.distributed import SlabGrid
    m, sch, c = _case("ks", 20000)
    g = SlabGrid(m, sch, c["x"], c["fields"], c["pars"], devices=[0])
    L, h = _lib.lib(), g.states[0].h
    assert L.tf_hook_set_dirichlet(h, 0, 1, 0.0, 0, 0.0) == _lib.TF_EINVAL      # periodic grid
    idt, nfs, le = C.c_double(), C.c_int(), C.c_double()
    assert L.tf_scheme_advance(h, sch.handle, 0.0, 0.1, 1e-2, 0.9, 100, 1e-12, 1, C.byref(idt),
                               C.byref(nfs), C.byref(le)) == _lib.TF_EINVAL
    g.close()
    out = C.c_void_p()
    cm = m._cuda
    # more ranks than tiles / a grid beyond the resident tiles of the given GPUs
    assert L.tf_state_create_slab(cm.ctx, cm.variant(()).handle, 3000, 1, 0, 8, C.byref(out)) == _lib.TF_EINVAL
    assert L.tf_state_create_slab(cm.ctx, cm.variant(()).handle, 1 << 23, 1, 0, 2, C.byref(out)) == _lib.TF_EINVAL
    assert L.tf_state_create_slab(cm.ctx, cm.variant(()).handle, 20000, 1, 2, 2, C.byref(out)) == _lib.TF_EINVAL


@pytest.mark.skipif(_ngpus() < 2, reason="needs two GPUs")
@pytest.mark.parametrize("name,N", CASES + [("ks", 1 << 21)])
def test_two_slabs_one_process(name, N):
    """Two GPUs driven by one process (peer pointers): halo values, scan records and the
    periodic border block cross the slab boundary inside the step kernel."""
    m, sch, c = _case(name, N)
    u, part, (tl, tt) = _slab(m, sch, c, 5, [0, 1])
    assert len(part) == 2 and part[1][0] == part[0][1] and sum(n for _, n in part) == N
    assert np.isfinite(u).all()
    assert rel_traj_err(u, _single(m, sch, c, 5)) <= 5e-10


@pytest.mark.skipif(_ngpus() < 2, reason="needs two GPUs")
def test_two_slabs_one_process_per_gpu():
    """torchrun, one process per GPU, record areas mapped through CUDA IPC handles."""
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node",
                          "2", "--master-addr", "127.0.0.1", "--master-port", "29517",
                          os.path.join(ROOT, "tools", "slab_check.py"), "dist"],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, (out.stdout + out.stderr)[-3000:]
    assert out.stdout.count(" ok ") >= 6 and "MISMATCH" not in out.stdout

"""CPU-side checks: the C-ABI library builds, loads and exports every symbol
declared in include/triflow_b200.h; it fails loudly without a GPU; host logic
(sharding, Dirichlet hook object, scheme tableaux)."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from triflow_b200 import _lib
    path = _lib.build_library()
    L = ctypes.CDLL(path)
    with open(os.path.join(ROOT, "include", "triflow_b200.h")) as f:
        header = f.read()
    declared = set(re.findall(r"\b(tf_[a-z_A-Z0-9]+)\s*\(", header))
    assert declared == set(_lib.EXPORTS)
    for name in declared:
        assert hasattr(L, name), name


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from triflow_b200 import _lib
    h = ctypes.c_void_p()
    rc = _lib.lib().tf_ctx_create(0, ctypes.byref(h))
    assert rc == _lib.TF_ECUDA
    assert b"no CPU fallback" in _lib.lib().tf_last_error()
    from triflow_b200 import workloads as W
    from triflow_b200.model import Model
    m = Model(**W.model_args("heat"), compiler="cuda")      # lowering needs no GPU
    x = np.linspace(0, 1, 50)
    with pytest.raises(_lib.CudaUnavailable):
        m.F(m.fields_template(x=x, T=x), dict(k=1, periodic=True))


def test_product_never_imports_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "triflow_b200")):
        for fn in files:
            if fn.endswith(".py"):
                src = open(os.path.join(dirpath, fn)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), fn


def test_shard_is_a_partition():
    from triflow_b200.distributed import shard
    for n, ws in [(32768, 8), (10, 3), (7, 8), (1, 1)]:
        cuts = [shard(n, r, ws) for r in range(ws)]
        assert cuts[0][0] == 0 and cuts[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(cuts, cuts[1:]))


def test_tableaux_match_oracle():
    from oracle import schemes as O
    from triflow_b200 import schemes as S
    for name in ("ROS2", "ROS3PRw", "ROS3PRL", "RODASPR"):
        a, g, b, bp = S.tableau(name)
        oa, og, ob, obp = O._build(name)
        assert np.array_equal(a, oa) and np.array_equal(g, og)
        assert list(b) == list(ob)
        assert (bp is None and obp is None) or list(bp) == list(obp)


def test_dirichlet_hook_is_a_plain_hook_too():
    from triflow_b200 import schemes as S
    from triflow_b200.fields import BaseFields
    f = BaseFields.factory1D(["U"], [])(x=np.arange(4.), U=np.zeros(4))
    f2, p = S.Dirichlet(U=(1, None))(0.0, f, {})
    assert f2["U"][0] == 1 and f2["U"][-1] == 0


def test_two_rank_gloo_shard_and_gather():
    """world_size 2 on CPU (gloo): sharded members gathered in member order."""
    script = r'''
import os, sys, numpy as np
sys.path.insert(0, %r)
from triflow_b200 import distributed as D
rank, ws = D.init("gloo")
lo, hi = D.shard(5, rank, ws)
local = np.arange(lo, hi, dtype=float)[:, None] * np.ones((1, 3))
D.barrier()
full = D.gather_members(local, 5)
mx = D.max_over_ranks(rank + 1.0)
sm = D.sum_over_ranks(rank + 1.0)
assert mx == 2.0 and sm == 3.0
if rank == 0:
    assert full.shape == (5, 3) and (full[:, 0] == np.arange(5)).all()
    print("GATHER_OK")
''' % ROOT
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29571")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1",
                          "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                          "--master-port", "29571", "-c", script] if False else
                         [sys.executable, "-c",
                          "import subprocess,sys,os;"
                          "ps=[subprocess.Popen([sys.executable,'-c',%r],env=dict(os.environ,"
                          "RANK=str(r),WORLD_SIZE='2',LOCAL_RANK=str(r)),stdout=subprocess.PIPE,"
                          "text=True) for r in range(2)];"
                          "outs=[p.communicate()[0] for p in ps];"
                          "print(''.join(outs)); sys.exit(max(p.returncode for p in ps))" % script],
                         env=env, capture_output=True, text=True, timeout=240)
    assert out.returncode == 0, out.stderr[-2000:]
    assert "GATHER_OK" in out.stdout


def test_slab_partition_covers_the_grid():
    """Slabs of whole tiles: consecutive, disjoint, cover [0, N); only the tail may be short."""
    from triflow_b200.distributed import slab_partition
    for N, nranks, tn, tl in ((2097152, 2, 3584, 293), (4194304, 4, 3584, 293), (20000, 2, 2048, 5),
                              (50001, 2, 2048, 13), (8388608, 8, 3584, 293)):
        part = slab_partition(N, nranks, tn, tl)
        assert len(part) == nranks and part[0][0] == 0
        for r in range(1, nranks):
            assert part[r][0] == part[r - 1][0] + part[r - 1][1] == r * tl * tn
        assert sum(n for _, n in part) == N and all(n > 0 for _, n in part)


def test_two_rank_gloo_slab_gather():
    """world_size 2 on CPU (gloo): the final gather of a slab grid (ragged last slab)."""
    script = r'''
import os, sys, numpy as np
sys.path.insert(0, %r)
from triflow_b200 import distributed as D
rank, ws = D.init("gloo")
part = D.slab_partition(50001, ws, 2048, 13)
off, n = part[rank]
full = D.gather_slabs(np.arange(off, off + n, dtype=float), part)
try:
    D.gather_slabs(np.zeros(3), part)
    raise SystemExit("a slab of the wrong size was accepted")
except ValueError:
    pass
if rank == 0:
    assert full.shape == (50001,) and (full == np.arange(50001)).all()
    print("SLAB_GATHER_OK")
''' % ROOT
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29573")
    out = subprocess.run([sys.executable, "-c",
                          "import subprocess,sys,os;"
                          "ps=[subprocess.Popen([sys.executable,'-c',%r],env=dict(os.environ,"
                          "RANK=str(r),WORLD_SIZE='2',LOCAL_RANK=str(r)),stdout=subprocess.PIPE,"
                          "text=True) for r in range(2)];"
                          "outs=[p.communicate()[0] for p in ps];"
                          "print(''.join(outs)); sys.exit(max(p.returncode for p in ps))" % script],
                         env=env, capture_output=True, text=True, timeout=240)
    assert out.returncode == 0, out.stderr[-2000:]
    assert "SLAB_GATHER_OK" in out.stdout


def test_host_pipeline_member_slicing():
    """HostPipeline splits fields / parameters along the member axis only."""
    from triflow_b200.ensemble import _members
    batch, N = 6, 10
    per_member = np.arange(batch, dtype=float)
    per_node = np.arange(N, dtype=float)
    full = np.arange(batch * N, dtype=float).reshape(batch, N)
    assert np.array_equal(_members(per_member, 2, 5, batch, N), per_member[2:5])
    assert _members(per_node, 2, 5, batch, N) is per_node
    assert np.array_equal(_members(full, 2, 5, batch, N), full[2:5])
    assert _members(0.25, 2, 5, batch, N) == 0.25
    # batch == N: fields are per node, parameters must say what they are
    sq = np.arange(N, dtype=float)
    assert _members(sq, 2, 5, N, N, field=True) is sq
    assert np.array_equal(_members(sq[:, None], 2, 5, N, N), sq[2:5, None])
    import pytest
    with pytest.raises(ValueError, match="ambiguous"):
        _members(sq, 2, 5, N, N)


def test_value_layout_classification():
    """1-D values: one per member or one per node; ambiguous when batch == N."""
    import pytest
    from triflow_b200.compiler import value_kind
    assert value_kind((), 4, 10) == "scalar"
    assert value_kind((4,), 4, 10) == "member" and value_kind((10,), 4, 10) == "node"
    assert value_kind((4, 10), 4, 10) == "member_node" and value_kind((4, 1), 4, 10) == "member"
    assert value_kind((1, 10), 4, 10) == "node" and value_kind((10,), 1, 10) == "node"
    assert value_kind((10, 1), 10, 10) == "member" and value_kind((1, 10), 10, 10) == "node"
    with pytest.raises(ValueError, match="ambiguous"):
        value_kind((10,), 10, 10)
    with pytest.raises(ValueError):
        value_kind((7,), 4, 10)

"""Host logic of the compiler plugin: uniform / per-node classification, the
host-evaluated constant table, parameter-shape handling, geometry helpers."""
import numpy as np
import pytest

from triflow_b200 import codegen, workloads as W
from triflow_b200.compiler import CompiledModel, default_chunk_nodes
from triflow_b200.model import Model


def lowered(name, node_pars=()):
    return codegen.lower(Model(**W.model_args(name), hold_compilation=True), node_pars)


def test_uniform_subtrees_are_hoisted_to_the_host():
    L = lowered("advdiff")
    # J of the linear model is entirely host-evaluated: no device arithmetic at all
    assert L.stats["J"]["ops"] == 0 and L.jacobian_is_constant
    assert "0.5 * c / dx + k / dx ** 2" in L.consts and "dx ** 2" in L.consts
    table = L.uniform_table(0.25, dict(k=2.0, c=3.0), 1)
    j = L.consts["0.5 * c / dx + k / dx ** 2"]
    assert table[0, j] == 0.5 * 3.0 / 0.25 + 2.0 / 0.25 ** 2
    assert table[0, L.n_const + j] == 1.0 / table[0, j]          # reciprocals for fast division


def test_per_system_parameters_give_one_row_per_member():
    L = lowered("advdiff")
    k = np.array([1.0, 2.0, 4.0])
    t = L.uniform_table(0.5, dict(k=k, c=1.0), 3)
    j = L.consts["dx ** 2"]
    assert np.all(t[:, j] == 0.25)
    jk = L.consts["0.5 * c / dx + k / dx ** 2"]
    assert np.array_equal(t[:, jk], 0.5 * 1.0 / 0.5 + k / 0.5 ** 2)


def test_array_parameter_changes_the_lowering():
    a, b = lowered("advdiff"), lowered("advdiff", ("k",))
    assert a.key != b.key and b.node_pars == ("k",) and not b.jacobian_is_constant
    assert "in.np[0]" in b.header and "in.np[0]" not in a.header
    assert "k" not in " ".join(b.consts)


def test_heaviside_is_one_and_max_min_are_numpy_semantics():
    L = lowered("burgers_up1")
    assert "Heaviside" not in L.header
    assert "TF_MAX(0.0, in.w[0][1])" in L.header and "TF_MIN(0.0, in.w[0][1])" in L.header


def test_j_scatter_tables_follow_the_reference_index_rule():
    L = lowered("film")                      # v=2, p=2: kk -> (eq, var, off)
    m = Model(**W.model_args("film"), hold_compilation=True)
    kk = np.asarray(m._sparse_indices[0])
    assert L.j_eq == (kk % 2).tolist()
    assert L.j_var == ((kk // 2) % 2).tolist()
    assert L.j_off == ((kk // 2) // 2 - 2).tolist()
    assert L.nnz == 10 and L.half_width == 2


def test_unsupported_function_is_reported():
    m = Model("k * dxxU + besselj(0, U)", "U", "k", hold_compilation=True)
    with pytest.raises(NotImplementedError):
        codegen.lower(m)


def test_parameter_shape_classification():
    m = Model(**W.model_args("advdiff"), hold_compilation=True)
    cm = CompiledModel(m)
    N = 64
    assert cm.node_pars_of(dict(k=1.0, c=2.0), N) == ()
    assert cm.node_pars_of(dict(k=np.ones(N), c=2.0), N) == ("k",)
    assert cm.node_pars_of(dict(k=np.ones(5), c=np.ones((5, N))), N, batch=5) == ("c",)
    with pytest.raises(ValueError):
        cm.node_pars_of(dict(k=np.ones(7), c=1.0), N)


def test_chunk_size_covers_the_band():
    for nvar, p in [(1, 1), (1, 2), (2, 1), (2, 2), (3, 1), (3, 2)]:
        mnodes = default_chunk_nodes(nvar, p)
        assert mnodes * nvar >= p * nvar + nvar - 1          # C >= BETA
        assert mnodes & (mnodes - 1) == 0


def test_generated_header_is_deterministic():
    assert lowered("ks").header == lowered("ks").header
    assert lowered("ks").key == lowered("ks").key

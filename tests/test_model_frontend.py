"""Host logic: the symbolic front-end reproduces the reference's expression
trees (golden: tests/golden/expr.json, dumped from the reference itself) and
the reference's API behaviour (reference tests/test_model.py:118-141,166-194)."""
import pickle

import pytest
import sympy as sp

from helpers import load_expr
from triflow_b200 import workloads as W
from triflow_b200.model import Model

EXPR = load_expr()


@pytest.mark.parametrize("name", sorted(W.MODELS))
def test_expression_trees_match_reference(name):
    m = Model(**W.model_args(name), hold_compilation=True)
    g = EXPR[name]
    assert [sp.srepr(e) for e in m.F_array.tolist()] == g["F"]
    assert [sp.srepr(e) for e in m.J_array.tolist()] == g["J"]
    assert [int(i) for i in m._sparse_indices[0]] == g["sparse_indices"]
    assert list(m._bounds) == g["bounds"]
    assert m._window_range == g["window_range"]
    assert m._nvar == g["nvar"]
    assert m._args == g["args"]


def test_api_errors_and_args():
    m = Model(["k * dxxU + s"], "U", "k", "s", hold_compilation=True)
    assert set(m._args) == {"x", "U_m1", "U", "U_p1", "s_m1", "s", "s_p1", "k", "dx"}
    with pytest.raises(NotImplementedError):
        Model("dxxxxxU", "U", hold_compilation=True)
    with pytest.raises(ValueError):
        Model("dxxx(dx)", "U", hold_compilation=True)
    with pytest.raises(NotImplementedError):
        Model("upwind(1, U, 4)", "U", hold_compilation=True)
    with pytest.raises(ValueError):
        Model("dxxU", "U", compiler="theano")


@pytest.mark.parametrize("spelling", ["k * dxxU", "k * dx(dxU)", ["k * dxxU"],
                                      "k * dx(U, 2)".replace("dx(U, 2)", "dxx(U)")])
def test_equation_spellings_agree(spelling):
    a = Model("k * dxxU", "U", "k", hold_compilation=True)
    b = Model(spelling, ["U"], ["k"], hold_compilation=True)
    assert a.F_array.tolist() == b.F_array.tolist()
    assert (a.J_array == b.J_array).all()


def test_pickle_roundtrip_keeps_definition():
    from oracle.numpy_compiler import numpy_compiler
    m = Model("k * dxxT", "T", "k", compiler=numpy_compiler)
    eqs, deps, pars, helps, bdcs = m.__reduce__()[1]
    m2 = Model(eqs, deps, pars, helps, bdcs, hold_compilation=True)
    assert m2.F_array.tolist() == m.F_array.tolist()
    assert (m2.J_array == m.J_array).all()
    assert m2._args == m._args
    assert pickle.dumps(m.fields_template(x=[0., 1.], T=[1., 2.]))

"""Translator check on the CPU: the generated F / J bodies, compiled with
g++ -ffp-contract=off against the host flavour of tf_model_prelude.h, reproduce
the reference's golden F and J (bit-exact where no libm pow is involved)."""
import ctypes
import os
import subprocess
import tempfile

import numpy as np
import pytest

from helpers import csc_triplet, fj_tags, load_fj, model_name_of
from triflow_b200 import codegen, workloads as W
from triflow_b200.model import Model

CSRC = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                    "triflow_b200", "csrc")

HARNESS = r"""
extern "C" void run(int N, int periodic, const double* fields, const double* npar,
                    const double* x, const double* cst, double* F, double* J) {
  for (int i = 0; i < N; ++i) {
    TfNodeIn in;
    for (int f = 0; f < TF_NFIELD; ++f)
      for (int o = 0; o < TF_WW; ++o) {
        int j = i + o - TF_P;
        if (periodic) j = ((j % N) + N) % N; else j = j < 0 ? 0 : (j >= N ? N - 1 : j);
        in.w[f][o] = fields[f * N + j];
      }
    for (int q = 0; q < TF_NNODEPAR; ++q) in.np[q] = npar[q * N + i];
    in.x = x[i];
    double f[TF_NVAR]; double jv[TF_NNZ];
    tf_model_F<TF_FAST_DIV != 0>(cst, in, f); tf_model_J<TF_FAST_DIV != 0>(cst, in, jv);
#if TF_F_SPLIT && defined(TF_TEST_SPLIT)
    tf_model_Fs<TF_FAST_DIV != 0>(cst, in, f);      // the monomial-collected form instead
#endif
    for (int e = 0; e < TF_NVAR; ++e) F[i * TF_NVAR + e] = f[e];
    for (int k = 0; k < TF_NNZ; ++k) J[i * TF_NNZ + k] = jv[k];
  }
}
"""

_CACHE = {}


def build(name, node_pars, fast_div, split=False):
    key = (name, node_pars, fast_div, split)
    if key in _CACHE:
        return _CACHE[key]
    m = Model(**W.model_args(name), hold_compilation=True)
    L = codegen.lower(m, node_pars)
    d = tempfile.mkdtemp(prefix="tfcg_")
    src = os.path.join(d, "m.cpp")
    with open(src, "w") as f:
        f.write(L.header + HARNESS)
    so = os.path.join(d, "m.so")
    subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-shared", "-fPIC",
                           "-DTF_FAST_DIV=%d" % fast_div, *(["-DTF_TEST_SPLIT"] if split else []),
                           "-I", CSRC, src, "-o", so])
    lib = ctypes.CDLL(so)
    _CACHE[key] = (m, L, lib)
    return _CACHE[key]


def run_case(tag, fast_div=0, split=False):
    x, fields, pars, F_ref, J_ref = load_fj(tag)
    name = model_name_of(tag)
    node_pars = tuple(k for k, v in pars.items()
                      if k != "periodic" and np.ndim(v) > 0)
    m, L, lib = build(name, node_pars, fast_div, split)
    if split and not L.f_split:
        pytest.skip("the model keeps the reference's form of F")
    N = x.size
    dx = (x[-1] - x[0]) / (N - 1)
    cst = L.uniform_table(dx, pars, 1)[0]
    fld = np.stack([fields[n] for n in L.fields]).astype(np.float64)
    npar = (np.stack([np.asarray(pars[q], float) for q in node_pars])
            if node_pars else np.zeros((1, N)))
    F = np.empty(N * L.nvar)
    J = np.empty((N, L.nnz))
    P = ctypes.POINTER(ctypes.c_double)
    as_p = lambda a: a.ctypes.data_as(P)
    lib.run(N, int(pars["periodic"]), as_p(fld), as_p(npar), as_p(x), as_p(cst),
            as_p(F), as_p(J))
    # assemble like the reference (COO -> CSC, duplicates summed in COO order)
    from scipy.sparse import csc_matrix
    i = np.arange(N)[:, None]
    eq, var, off = (np.array(a)[None, :] for a in (L.j_eq, L.j_var, L.j_off))
    j = i + off
    j = j % N if pars["periodic"] else np.clip(j, 0, N - 1)
    Jm = csc_matrix((J.reshape(-1), ((i * L.nvar + eq).reshape(-1),
                                     (j * L.nvar + var).reshape(-1))),
                    shape=(N * L.nvar, N * L.nvar))
    return F, Jm, F_ref, J_ref


@pytest.mark.parametrize("tag", fj_tags())
def test_generated_code_matches_reference(tag):
    F, J, F_ref, J_ref = run_case(tag)
    uses_pow = "film" in tag                      # h**3 goes through libm pow
    if uses_pow:
        assert np.max(np.abs(F - F_ref)) <= 1e-15 * np.max(np.abs(F_ref))
    else:
        assert np.array_equal(F, F_ref)
    ip, ix, dat = csc_triplet(J)
    ipr, ixr, datr = csc_triplet(J_ref)
    assert np.array_equal(ip, ipr) and np.array_equal(ix, ixr)
    if uses_pow:
        assert np.max(np.abs(dat - datr)) <= 4e-16 * np.max(np.abs(datr))
    else:
        assert np.array_equal(dat, datr)


@pytest.mark.parametrize("tag", ["ks_per", "advdiff_edge", "film_per", "burgers_up2_per"])
def test_fast_division_mode_within_one_ulp(tag):
    F, J, F_ref, J_ref = run_case(tag, fast_div=1)
    assert np.max(np.abs(F - F_ref)) <= 1e-13 * np.max(np.abs(F_ref))
    assert abs(J - J_ref).max() <= 1e-13 * abs(J_ref).max()


@pytest.mark.parametrize("tag", [t for t in fj_tags() if model_name_of(t) in (
    "ks", "burgers_up1", "burgers_up2", "burgers_up3", "kdv", "burgers_central", "helper_dx")])
def test_monomial_collected_form_of_F(tag):
    """The form of F the solver kernels use for nonlinear polynomial-like models (one collected,
    host-evaluated coefficient per monomial of the stencil values) against the reference's golden
    F: the same function, different rounding.  The bound is the rounding of the terms that
    cancel: eps x sum of |monomial terms|, far above |F| on fine grids."""
    F, J, F_ref, J_ref = run_case(tag, fast_div=1, split=True)
    x, fields, pars, _, _ = load_fj(tag)
    dx = (x[-1] - x[0]) / (x.size - 1)
    p = 3 if "up3" in tag else 2
    mag = max(np.max(np.abs(v)) for v in fields.values())
    # largest single term of the stencil polynomial: |coefficient| x |u| (x |u| for the products)
    big = max(1.0, mag) * mag * max(1.0, 6.0 / dx ** 4 if "ks" in tag else 4.0 / dx ** 3)
    assert np.max(np.abs(F - F_ref)) <= 64 * np.finfo(float).eps * big
    assert np.max(np.abs(F - F_ref)) <= 1e-9 * np.max(np.abs(F_ref))


def test_linear_models_are_recognised():
    """F == sum_k J_k u_k is decided symbolically; only then may the solver kernels use the sum."""
    from triflow_b200 import codegen, workloads as W
    from triflow_b200.model import Model
    want = {"advdiff": True, "heat": True, "coupled": True, "helper": False, "ks": False,
            "burgers_up1": False, "film": False}
    for name, lin in want.items():
        L = codegen.lower(Model(**W.model_args(name), hold_compilation=True))
        assert L.f_is_linear is lin, name
        assert ("#define TF_F_LINEAR %d" % int(lin)) in L.header
    # per-node coefficients: J is not a table of constants any more -> general form
    m = Model(**W.model_args("advdiff"), hold_compilation=True)
    assert codegen.lower(m, ("k",)).f_is_linear is False
    # an affine model (constant source term) is not homogeneous linear
    src = Model("k * dxxU - c * dxU + 1", "U", ["k", "c"], hold_compilation=True)
    assert codegen.lower(src).f_is_linear is False


def test_which_models_get_the_monomial_collected_form():
    from triflow_b200 import codegen, workloads as W
    from triflow_b200.model import Model
    want = {"ks": True, "burgers_up1": True, "kdv": True, "advdiff": False, "heat": False,
            "film": False, "helper": False}
    for name, split in want.items():
        L = codegen.lower(Model(**W.model_args(name), hold_compilation=True))
        assert L.f_split is split, name
        assert ("tf_model_Fs" in L.header) is split
        if split:                                  # no division left on the device, fewer operations
            assert L.stats["Fs"]["div"] == 0 and L.stats["Fs"]["divc"] == 0
            assert L.stats["Fs"]["ops"] < L.stats["F"]["ops"] + 3 * L.stats["F"]["divc"]
    # per-node coefficients keep the reference's form
    assert codegen.lower(Model(**W.model_args("ks"), hold_compilation=True), ()).f_split

"""Pin the CPU oracle against fixtures generated from the reference itself
(tests/golden/make_golden.py).  F / J: bit-exact.  Trajectories: 1e-13 (they go
through SuperLU)."""
import numpy as np
import pytest

from helpers import csc_triplet, fj_tags, load_fj, model_name_of, rel_traj_err, traj
from oracle import schemes as O
from oracle.numpy_compiler import numpy_compiler
from triflow_b200 import workloads as W
from triflow_b200.model import Model

_MODELS = {}


def omodel(name):
    if name not in _MODELS:
        _MODELS[name] = Model(**W.model_args(name), compiler=numpy_compiler)
    return _MODELS[name]


@pytest.mark.parametrize("tag", fj_tags())
def test_F_J_bit_exact(tag):
    x, fields, pars, F_ref, J_ref = load_fj(tag)
    m = omodel(model_name_of(tag))
    f = m.fields_template(x=x, **fields)
    F = m.F(f, pars)
    assert F.dtype == np.float64 and np.array_equal(F, F_ref)
    ip, ix, dat = csc_triplet(m.J(f, pars))
    ipr, ixr, datr = csc_triplet(J_ref)
    assert np.array_equal(ip, ipr) and np.array_equal(ix, ixr)
    assert np.array_equal(dat, datr)


def run_fixed(m, scheme, c, steps, every, hook=None, pars=None):
    f = m.fields_template(x=c["x"], **c["fields"])
    pars = c["pars"] if pars is None else pars
    t, snaps = 0.0, []
    for i in range(steps):
        t, f = scheme(t, f, c["dt"], pars, hook=hook or O.null_hook)
        if (i + 1) % every == 0:
            snaps.append(f.uflat.copy())
    return np.array(snaps)


TOL = 1e-13


@pytest.mark.parametrize("sname,kw", [
    ("ROS3PRw", dict(time_stepping=False)), ("ROS2", {}), ("Theta", dict(theta=1)),
    ("Theta05", dict(theta=.5)), ("ROS3PRL", dict(time_stepping=False)),
    ("RODASPR", dict(time_stepping=False))])
def test_readme_fixed(sname, kw):
    g = traj()
    c = W.readme(200)
    m = omodel("advdiff")
    cls = getattr(O, "Theta" if sname.startswith("Theta") else sname)
    snaps = run_fixed(m, cls(m, **kw), c, 5, 1, hook=W.readme_hook)
    assert rel_traj_err(snaps, g["readme_fixed_" + sname]) <= TOL


def test_readme_adaptive_controller_trace():
    g = traj()
    c = W.readme(200)
    m = omodel("advdiff")
    sch = O.ROS3PRw(m, tol=1e-1)
    f = m.fields_template(x=c["x"], **c["fields"])
    t, snaps, counts = 0.0, [], []
    for _ in range(5):
        n0 = sch.n_fixed_steps
        f, _p = W.readme_hook(t, f, c["pars"])
        t, f = sch(t, f, c["dt"], c["pars"], hook=W.readme_hook)
        snaps.append(f.uflat.copy())
        counts.append(sch.n_fixed_steps - n0)
    assert counts == list(g["readme_adaptive_counts"]) == [55, 10, 13, 14, 10]
    assert rel_traj_err(np.array(snaps), g["readme_adaptive_ROS3PRw"]) <= 1e-11
    assert abs(snaps[-1].sum() - 16.7597312006418) < 1e-9


def test_readme_simulation_default_is_double_wrapped():
    from triflow_b200.simulation import Simulation
    g = traj()
    c = W.readme(200)
    m = omodel("advdiff")
    sim = Simulation(m, dict(x=c["x"], **c["fields"]), c["pars"], dt=c["dt"],
                     tmax=c["tmax"], hook=W.readme_hook, scheme=O.ROS3PRw)
    snaps = np.array([f.uflat.copy() for _, f in sim])
    assert rel_traj_err(snaps, g["readme_simdefault_ROS3PRw"]) <= 1e-10
    sim = Simulation(m, dict(x=c["x"], **c["fields"]), c["pars"], dt=c["dt"],
                     tmax=c["tmax"], hook=W.readme_hook, scheme=O.ROS3PRw,
                     time_stepping=False)
    snaps = np.array([f.uflat.copy() for _, f in sim])
    assert rel_traj_err(snaps, g["readme_simfixed_ROS3PRw"]) <= TOL


@pytest.mark.parametrize("acc", [1, 2])
def test_burgers(acc):
    c = W.burgers(2048, acc)
    m = omodel(c["model"])
    snaps = run_fixed(m, O.ROS2(m), c, 50, 10)
    assert rel_traj_err(snaps, traj()["burgers_up%d_2048" % acc]) <= TOL


@pytest.mark.parametrize("N", [2048, 1000])
def test_ks(N):
    c = W.kuramoto(N)
    m = omodel("ks")
    snaps = run_fixed(m, O.ROS3PRw(m, time_stepping=False), c, 50, 10)
    assert rel_traj_err(snaps, traj()["ks_%d" % N]) <= 1e-11   # chaotic growth


def test_ks_edge():
    c = W.kuramoto(512)
    m = omodel("ks")
    snaps = run_fixed(m, O.ROS3PRw(m, time_stepping=False), c, 20, 5,
                      pars=dict(periodic=False))
    assert rel_traj_err(snaps, traj()["ks_512_edge"]) <= 1e-11


@pytest.mark.parametrize("theta", [1, .5])
def test_film(theta):
    c = W.film(1024, theta)
    m = omodel("film")
    snaps = run_fixed(m, O.Theta(m, theta=theta), c, 100, 20)
    assert rel_traj_err(snaps, traj()["film_1024_theta%g" % theta]) <= 1e-11


def test_ensemble_members():
    g = traj()
    mem = g["ensemble_512_members"]
    c = W.ensemble(512, mem)
    m = omodel("advdiff")
    for idx in range(0, len(mem), 5):
        pars = dict(k=float(c["pars"]["k"][idx]), c=float(c["pars"]["c"][idx]),
                    periodic=False)
        snaps = run_fixed(m, O.ROS3PRw(m, time_stepping=False), c, 100, 100,
                          hook=W.readme_hook, pars=pars)
        assert rel_traj_err(snaps[-1], g["ensemble_512_final"][idx]) <= TOL


def test_ensemble_members_full_grid():
    """cfg 5 at N = 4096: the oracle against three of the 64 members the reference ran."""
    from helpers import traj_ens4096
    g = traj_ens4096()
    mem = g["members"]
    c = W.ensemble(4096, mem)
    m = omodel("advdiff")
    for idx in (0, len(mem) // 2, len(mem) - 1):
        pars = dict(k=float(c["pars"]["k"][idx]), c=float(c["pars"]["c"][idx]),
                    periodic=False)
        snaps = run_fixed(m, O.ROS3PRw(m, time_stepping=False), c, 100, 100,
                          hook=W.readme_hook, pars=pars)
        assert rel_traj_err(snaps[-1], g["final"][idx]) <= TOL


@pytest.mark.parametrize("sname,kw", [("ROS2", {}),
                                      ("ROS3PRw", dict(time_stepping=False)),
                                      ("Theta", {})])
def test_heat50(sname, kw):
    x = np.linspace(0, 10, 50, endpoint=False)
    c = dict(x=x, fields=dict(T=np.cos(x * 2 * np.pi / 10)),
             pars=dict(k=1, periodic=True), dt=1.0)
    m = omodel("heat")
    snaps = run_fixed(m, getattr(O, sname)(m, **kw), c, 20, 5)
    assert rel_traj_err(snaps, traj()["heat50_" + sname]) <= TOL

"""Drop-in tests: the REFERENCE's own, unmodified objects drive this repository's plug-ins.

``reference.Model(..., compiler=cuda_compiler)`` (core/model.py:145-155,299-311) builds the
reference's F_Routine / J_Routine around the CUDA callables; ``reference.Simulation(model, ...,
scheme=triflow_b200.schemes.ROS3PRw)`` (core/simulation.py:160-261) constructs and iterates
the device scheme exactly as it does with its own.  Compared with the fixtures the reference
produced with its numpy compiler + SciPy path (tests/golden).

The reference sources are loaded from ``baseline/_ref`` (copied by ``build()``; travels to the
GPU box) -- see oracle/ref_loader.py for the import shims."""
import numpy as np
import pytest

from oracle import ref_loader
from helpers import csc_triplet, load_fj, rel_traj_err, traj

needs_ref = pytest.mark.skipif(ref_loader.reference_root() is None,
                               reason="reference sources not available (run build() first)")


def ref_model(name, compiler):
    from triflow_b200 import workloads as W
    a = W.model_args(name)
    Model = ref_loader.load_reference()["model"].Model
    return Model(a["differential_equations"], a["dependent_variables"], a["parameters"],
                 a["help_functions"], compiler=compiler)


@needs_ref
def test_reference_loader_runs_the_reference_numpy_path():
    """(CPU) the loader really executes the reference: its numpy compiler reproduces a golden
    F / J pair bit for bit."""
    x, fields, pars, F_ref, J_ref = load_fj("ks_per")
    m = ref_model("ks", "numpy")
    f = m.fields_template(x=x, **fields)
    assert np.array_equal(m.F(f, pars), F_ref)
    for a, b in zip(csc_triplet(m.J(f, pars)), csc_triplet(J_ref)):
        assert np.array_equal(a, b)


@needs_ref
@pytest.mark.gpu
@pytest.mark.parametrize("tag", ["advdiff_edge", "ks_per", "burgers_up2_per", "coupled_edge"])
def test_reference_model_with_cuda_compiler(tag):
    """reference.Model + reference F_Routine / J_Routine around the CUDA compiler plugin:
    bit-exact F and J (incl. the reference's own diff_approx helper running on top)."""
    from helpers import model_name_of
    from triflow_b200.compiler import cuda_compiler
    x, fields, pars, F_ref, J_ref = load_fj(tag)
    m = ref_model(model_name_of(tag), cuda_compiler)
    assert type(m).__module__ == "triflow.core.model"
    f = m.fields_template(x=x, **fields)
    assert np.array_equal(m.F(f, pars), F_ref)
    ip, ix, dat = csc_triplet(m.J(f, pars))
    ipr, ixr, datr = csc_triplet(J_ref)
    assert np.array_equal(ip, ipr) and np.array_equal(ix, ixr) and np.array_equal(dat, datr)
    assert np.array_equal(np.asarray(m.J(f, pars, sparse=False)), J_ref.toarray())


@needs_ref
@pytest.mark.gpu
def test_reference_simulation_drives_the_device_schemes():
    """reference.Simulation iterating reference.Model(compiler=cuda_compiler) with
    triflow_b200.schemes.ROS3PRw: fixed-step and the reference's default (its Richardson
    wrapper around our scheme object, simulation.py:190-197) against the reference's own run."""
    from triflow_b200 import schemes as S, workloads as W
    from triflow_b200.compiler import cuda_compiler
    mods = ref_loader.load_reference()
    Sim = mods["simulation"].Simulation
    g = traj()
    c = W.readme(200)
    m = ref_model("advdiff", cuda_compiler)
    sim = Sim(m, dict(x=c["x"], **c["fields"]), c["pars"], dt=c["dt"], tmax=c["tmax"],
              hook=W.readme_hook, scheme=S.ROS3PRw, time_stepping=False)
    snaps = np.array([f.uflat.copy() for _, f in sim])
    assert sim.t == 2.5          # (the reference sets `status` in run(), not when iterated)
    assert rel_traj_err(snaps, g["readme_simfixed_ROS3PRw"]) <= 1e-8
    assert abs(snaps[-1].sum() - 13.293911092223974) < 1e-8      # SURVEY.md §8c smoke value
    # declarative hook: stays on the device
    sim = Sim(m, dict(x=c["x"], **c["fields"]), c["pars"], dt=c["dt"], tmax=c["tmax"],
              hook=S.Dirichlet(U=(1, 0)), scheme=S.ROS3PRw, time_stepping=False)
    snaps = np.array([f.uflat.copy() for _, f in sim])
    assert rel_traj_err(snaps, g["readme_simfixed_ROS3PRw"]) <= 1e-8
    # the reference's default: its own time_stepping wrapper around our scheme
    sim = Sim(m, dict(x=c["x"], **c["fields"]), c["pars"], dt=c["dt"], tmax=c["tmax"],
              hook=W.readme_hook, scheme=S.ROS3PRw)
    snaps = np.array([f.uflat.copy() for _, f in sim])
    assert rel_traj_err(snaps, g["readme_simdefault_ROS3PRw"]) <= SIMDEFAULT_TOL


# measured on B200 (tools/sim_default_err.py): 2e-16 ... 3e-16
SIMDEFAULT_TOL = 1e-8


@needs_ref
@pytest.mark.gpu
def test_reference_schemes_on_the_cuda_compiler():
    """The other direction: the reference's OWN ROS3PRw / Theta (SciPy SuperLU) stepping a
    model whose F / J come from the CUDA compiler plugin -> identical to its numpy path."""
    from triflow_b200 import workloads as W
    from triflow_b200.compiler import cuda_compiler
    mods = ref_loader.load_reference()
    g = traj()
    c = W.readme(200)
    m = ref_model("advdiff", cuda_compiler)
    sch = mods["schemes"].ROS3PRw(m, time_stepping=False)
    f = m.fields_template(x=c["x"], **c["fields"])
    t, snaps = 0.0, []
    for _ in range(5):
        t, f = sch(t, f, c["dt"], c["pars"], hook=W.readme_hook)
        snaps.append(f.uflat.copy())
    assert np.array_equal(np.array(snaps), g["readme_fixed_ROS3PRw"])


@needs_ref
@pytest.mark.gpu
def test_interpolated_output_mode_equals_the_reference():
    """recompute_target=False (schemes.py:183-187,217-222): the embedded controller overshoots
    the output time and the output is interpolated between the last two internal states, the
    interpolant serving later output times too.  Reference (numpy compiler + SuperLU) against
    this package's ROS3PRw on the README problem, same hook, five outputs."""
    from triflow_b200 import schemes as S, workloads as W
    from triflow_b200.model import Model
    mods = ref_loader.load_reference()
    c = W.readme(200)
    rm = ref_model("advdiff", "numpy")
    rs = mods["schemes"].ROS3PRw(rm, tol=1e-1, recompute_target=False)
    gm = Model(**W.model_args("advdiff"), compiler="cuda")
    gs = S.ROS3PRw(gm, tol=1e-1, recompute_target=False)
    rf = rm.fields_template(x=c["x"], **c["fields"])
    gf = gm.fields_template(x=c["x"], **c["fields"])
    tr = tg = 0.0
    for _ in range(5):
        tr, rf = rs(tr, rf, c["dt"], c["pars"], hook=W.readme_hook)
        tg, gf = gs(tg, gf, c["dt"], c["pars"], hook=W.readme_hook)
        assert tr == tg
        assert rel_traj_err(gf.uflat, rf.uflat) <= 1e-8
        assert abs(gs._internal_dt - rs._internal_dt) <= 1e-9 * abs(rs._internal_dt)


@needs_ref
@pytest.mark.gpu
def test_theta_with_a_user_solver_equals_the_reference():
    """Theta(model, theta, solver=callable) (schemes.py:518-521,548-559): the user's solver is
    called on the host with the same A and b as in the reference (F, J from the device)."""
    import scipy.sparse.linalg as spl
    from triflow_b200 import schemes as S, workloads as W
    from triflow_b200.model import Model
    mods = ref_loader.load_reference()
    calls = []

    def solver(A, b):
        calls.append(A.shape)
        return spl.spsolve(A, b)
    c = W.readme(200)
    rm = ref_model("advdiff", "numpy")
    gm = Model(**W.model_args("advdiff"), compiler="cuda")
    for theta in (1, 0.5, 0):
        rs = mods["schemes"].Theta(rm, theta=theta, solver=solver)
        gs = S.Theta(gm, theta=theta, solver=solver)
        rf = rm.fields_template(x=c["x"], **c["fields"])
        gf = gm.fields_template(x=c["x"], **c["fields"])
        tr = tg = 0.0
        for _ in range(3):
            tr, rf = rs(tr, rf, 1e-3, c["pars"], hook=W.readme_hook)
            tg, gf = gs(tg, gf, 1e-3, c["pars"], hook=W.readme_hook)
        assert np.array_equal(gf.uflat, rf.uflat)          # bit-identical F, J -> same A, b
    assert len(calls) == 18 and calls[0] == (200, 200)
    # without a solver the device banded solver does the same step
    gd = S.Theta(gm, theta=0.5)
    gf = gm.fields_template(x=c["x"], **c["fields"])
    rf = rm.fields_template(x=c["x"], **c["fields"])
    rs = mods["schemes"].Theta(rm, theta=0.5)
    tg, gf = gd(0.0, gf, 1e-3, c["pars"], hook=W.readme_hook)
    tr, rf = rs(0.0, rf, 1e-3, c["pars"], hook=W.readme_hook)
    assert rel_traj_err(gf.uflat, rf.uflat) <= 1e-8

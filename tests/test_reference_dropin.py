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

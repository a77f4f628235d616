"""The five BASELINE.json workloads as concrete synthetic inputs.

Shared by ``bench.py``, the parity tests and ``tests/golden/make_golden.py`` so
that the GPU path, the CPU oracle and the reference itself all see the same
model strings, grids, initial conditions and parameters.  Every generator takes
the node count ``N`` so the same workload can be instantiated at CPU-oracle
sizes and at the full BASELINE sizes.
"""

import numpy as np

FILM_EQS = ["-dxq",
            "(5/6*h - 5/2*q/h**2 - 17/7*q/h*dxq + (9/7*q**2/h**2 - 5/6*B*h)*dxh"
            " + 5/6*G*h*dxxxh + 4*q/h**2*dxh**2 - 9/2/h*dxq*dxh - 6*q/h*dxxh"
            " + 9/2*dxxq)/delta"]

MODELS = {
    # name: (differential_equations, dependent_variables, parameters, help_functions)
    "advdiff": ("k * dxxU - c * dxU", "U", ["k", "c"], None),
    "heat": ("k * dxxT", "T", "k", None),
    "burgers_up1": ("k*dxxU - upwind(U, U, 1)", "U", "k", None),
    "burgers_up2": ("k*dxxU - upwind(U, U, 2)", "U", "k", None),
    "burgers_up3": ("k*dxxU - upwind(U, U, 3)", "U", "k", None),
    "burgers_central": ("k*dxxU - U*dxU", "U", "k", None),
    "ks": ("-dxxU - dxxxxU + (dxU)**2", "U", None, None),
    "film": (FILM_EQS, ["h", "q"], ["delta", "B", "G"], None),
    "bivariate": (["k1 * dxx(v)", "k2 * dxx(u)"], ["u", "v"], ["k1", "k2"], None),
    "coupled": (["k1 * dxxU - c1 * dxV", "k2 * dxxV - c2 * dxU"], ["U", "V"],
                ["k1", "k2", "c1", "c2"], None),
    "helper": (["k * dxxU + s"], "U", "k", "s"),
    "helper_dx": (["k * dxxU - dxs * U"], "U", "k", "s"),
    "upwind_const": (["upwind(1, U, 2)"], "U", "k", "s"),
    "kdv": ("-c*dxU - dxxxU*b - U*dxU", "U", ["c", "b"], None),
}


def model_args(name):
    eqs, deps, pars, helps = MODELS[name]
    return dict(differential_equations=eqs, dependent_variables=deps,
                parameters=pars, help_functions=helps)


# --------------------------------------------------------------- cfg 1 (README)
def readme(N=200):
    x = np.linspace(0, 1, N)
    U = np.cos(2 * np.pi * x * 5)
    pars = dict(c=.03, k=.001, periodic=False)
    return dict(model="advdiff", x=x, fields=dict(U=U), pars=pars, scheme="ROS3PRw",
                dt=0.5, tmax=2.5, dirichlet={"U": (1.0, 0.0)})


def readme_hook(t, fields, pars):
    fields["U"][0] = 1
    fields["U"][-1] = 0
    return fields, pars


# -------------------------------------------------------------- cfg 2 (Burgers)
def burgers(N=2 ** 17, accuracy=1):
    dx = 0.2
    x = np.arange(N) * dx
    # both wavelengths divide L = N*dx for N a multiple of 512
    U = np.sin(2 * np.pi * x / 102.4) + 0.5 * np.sin(2 * np.pi * x / 40.96 + 1)
    return dict(model="burgers_up%d" % accuracy, x=x, fields=dict(U=U),
                pars=dict(k=0.1, periodic=True), scheme="ROS2", dt=0.1,
                parity_steps=50, bench_steps=200)


# ------------------------------------------------------------------- cfg 3 (KS)
def kuramoto(N=2 ** 20, seed=0):
    dx = 200 / 2009
    x = np.arange(N) * dx
    L = N * dx
    m = max(1, round(L / 20))
    rng = np.random.default_rng(seed)
    U = 2 * np.cos(2 * np.pi * m * x / L) + 5 + 1e-3 * rng.standard_normal(N)
    return dict(model="ks", x=x, fields=dict(U=U), pars=dict(periodic=True),
                scheme="ROS3PRw", dt=0.2, parity_steps=50, bench_steps=100)


# ----------------------------------------------------------------- cfg 4 (film)
def film(N=2 ** 18, theta=1):
    dx = 0.25
    x = np.arange(N) * dx
    h = 1 + 0.1 * np.cos(2 * np.pi * x / 64)
    q = h ** 3 / 3
    return dict(model="film", x=x, fields=dict(h=h, q=q),
                pars=dict(delta=10, B=0.1, G=50, periodic=True), scheme="Theta",
                theta=theta, dt=0.05, parity_steps=100, bench_steps=200)


# ------------------------------------------------------------- cfg 5 (ensemble)
ENSEMBLE_K = 256
ENSEMBLE_C = 128


def ensemble_member(r):
    """(k, c) of member ``r = i*128 + j``."""
    i, j = divmod(int(r), ENSEMBLE_C)
    k = np.geomspace(2.5e-4, 4e-3, ENSEMBLE_K)[i]
    c = np.linspace(-0.06, 0.06, ENSEMBLE_C)[j]
    return k, c


def ensemble(N=4096, members=None):
    members = np.arange(ENSEMBLE_K * ENSEMBLE_C) if members is None \
        else np.asarray(members)
    x = np.linspace(0, 1, N)
    U = np.cos(2 * np.pi * x * 5)
    i, j = np.divmod(members, ENSEMBLE_C)
    k = np.geomspace(2.5e-4, 4e-3, ENSEMBLE_K)[i]
    c = np.linspace(-0.06, 0.06, ENSEMBLE_C)[j]
    return dict(model="advdiff", x=x, fields=dict(U=U), members=members,
                pars=dict(k=k, c=c, periodic=False), scheme="ROS3PRw", dt=0.025,
                parity_steps=100, bench_steps=100, dirichlet={"U": (1.0, 0.0)})


def ensemble_parity_subset(n_random=59, seed=0):
    rng = np.random.default_rng(seed)
    fixed = [0, 127, 128, 16384, 32767]
    rnd = rng.choice(ENSEMBLE_K * ENSEMBLE_C, size=n_random, replace=False)
    return np.array(sorted(set(fixed) | set(int(r) for r in rnd)))

#!/usr/bin/env python
"""GPU bring-up script: runs every parity case and PRINTS the error magnitudes
(continues past failures) so that one gpurun call gives the whole picture."""
import os
import sys
import time
import traceback

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from helpers import csc_triplet, fj_tags, load_fj, model_name_of, rel_traj_err, traj  # noqa
from oracle import schemes as O  # noqa
from oracle.numpy_compiler import numpy_compiler  # noqa
from triflow_b200 import schemes as S, workloads as W  # noqa
from triflow_b200.model import Model  # noqa

GM, OM = {}, {}


def gmodel(name):
    if name not in GM:
        GM[name] = Model(**W.model_args(name), compiler="cuda")
    return GM[name]


def omodel(name):
    if name not in OM:
        OM[name] = Model(**W.model_args(name), compiler=numpy_compiler)
    return OM[name]


def section(title):
    print("\n==== %s" % title, flush=True)


def attempt(label, fn):
    t0 = time.time()
    try:
        r = fn()
        print("%-46s %s  (%.1fs)" % (label, r, time.time() - t0), flush=True)
    except Exception as e:  # noqa
        print("%-46s EXC %s: %s" % (label, type(e).__name__, str(e)[:300]), flush=True)
        if os.environ.get("TF_TRACE"):
            traceback.print_exc()


def fj(tag):
    x, fields, pars, F_ref, J_ref = load_fj(tag)
    m = gmodel(model_name_of(tag))
    f = m.fields_template(x=x, **fields)
    F = m.F(f, pars)
    J = m.J(f, pars)
    dF = np.max(np.abs(F - F_ref)) / max(np.max(np.abs(F_ref)), 1e-300)
    ip, ix, dat = csc_triplet(J)
    ipr, ixr, datr = csc_triplet(J_ref)
    same = np.array_equal(ip, ipr) and np.array_equal(ix, ixr)
    dJ = (np.max(np.abs(dat - datr)) / np.max(np.abs(datr))) if same else float("nan")
    return "F bit=%s rel=%.1e | J struct=%s bit=%s rel=%.1e" % (
        np.array_equal(F, F_ref), dF, same, same and np.array_equal(dat, datr), dJ)


def run_fixed(m, scheme, c, steps, every, hook=None, pars=None):
    f = m.fields_template(x=c["x"], **c["fields"])
    pars = c["pars"] if pars is None else pars
    t, snaps = 0.0, []
    for i in range(steps):
        t, f = scheme(t, f, c["dt"], pars, hook=hook or S.null_hook)
        if (i + 1) % every == 0:
            snaps.append(f.uflat.copy())
    return np.array(snaps)


def main():
    g = traj()
    section("F / J through the plugin ufunc contract vs reference golden")
    for tag in fj_tags():
        attempt(tag, lambda tag=tag: fj(tag))

    section("one step vs oracle (solver path)")

    def one_step(name, cfg, scls, ocls, kw, hook=None, pars=None):
        gm, om = gmodel(name), omodel(name)
        pars = cfg["pars"] if pars is None else pars
        f0 = gm.fields_template(x=cfg["x"], **cfg["fields"])
        _, fg = scls(gm, **kw)(0.0, f0, cfg["dt"], pars, hook=hook or S.null_hook)
        _, fo = ocls(om, **kw)(0.0, f0, cfg["dt"], pars, hook=hook or O.null_hook)
        return "rel err %.2e" % rel_traj_err(fg.uflat, fo.uflat)

    fx = dict(time_stepping=False)
    attempt("heat N=50 per ROS2", lambda: one_step(
        "heat", dict(x=np.linspace(0, 10, 50, endpoint=False),
                     fields=dict(T=np.cos(np.linspace(0, 10, 50, endpoint=False) * 2 * np.pi / 10)),
                     pars=dict(k=1, periodic=True), dt=1.0), S.ROS2, O.ROS2, {}))
    attempt("advdiff N=200 edge ROS3PRw dirichlet", lambda: one_step(
        "advdiff", W.readme(200), S.ROS3PRw, O.ROS3PRw, fx, hook=S.Dirichlet(U=(1, 0))))
    attempt("advdiff N=200 per ROS3PRw", lambda: one_step(
        "advdiff", W.readme(200), S.ROS3PRw, O.ROS3PRw, fx,
        pars=dict(W.readme(200)["pars"], periodic=True)))
    for N in (256, 300, 2048, 5000, 70000):
        attempt("ks N=%d per ROS3PRw" % N, lambda N=N: one_step(
            "ks", W.kuramoto(N), S.ROS3PRw, O.ROS3PRw, fx))
    attempt("ks N=512 edge ROS3PRw", lambda: one_step(
        "ks", W.kuramoto(512), S.ROS3PRw, O.ROS3PRw, fx, pars=dict(periodic=False)))
    for acc in (1, 2):
        attempt("burgers_up%d N=2048 ROS2" % acc, lambda acc=acc: one_step(
            "burgers_up%d" % acc, W.burgers(2048, acc), S.ROS2, O.ROS2, {}))
    attempt("burgers_up1 N=40000 ROS2", lambda: one_step(
        "burgers_up1", W.burgers(40000 // 512 * 512, 1), S.ROS2, O.ROS2, {}))
    if not os.environ.get("TF_SKIP_FILM"):
        attempt("film N=1024 Theta", lambda: one_step(
            "film", W.film(1024), S.Theta, O.Theta, {}))

    section("trajectories vs reference golden")
    c = W.readme(200)
    for sname, kw in [("ROS3PRw", fx), ("ROS2", {}), ("Theta", dict(theta=1)),
                      ("Theta05", dict(theta=.5)), ("ROS3PRL", fx), ("RODASPR", fx)]:
        cls = getattr(S, "Theta" if sname.startswith("Theta") else sname)
        attempt("readme fixed %s (Dirichlet hook)" % sname, lambda cls=cls, kw=kw, sname=sname:
                "%.2e" % rel_traj_err(run_fixed(gmodel("advdiff"), cls(gmodel("advdiff"), **kw), c, 5, 1,
                                                hook=S.Dirichlet(U=(1, 0))), g["readme_fixed_" + sname]))
    attempt("readme fixed ROS3PRw (python hook)", lambda: "%.2e" % rel_traj_err(
        run_fixed(gmodel("advdiff"), S.ROS3PRw(gmodel("advdiff"), **fx), c, 5, 1, hook=W.readme_hook),
        g["readme_fixed_ROS3PRw"]))

    def adaptive(hook):
        m = gmodel("advdiff")
        sch = S.ROS3PRw(m, tol=1e-1)
        f = m.fields_template(x=c["x"], **c["fields"])
        t, snaps, counts = 0.0, [], []
        for _ in range(5):
            n0 = sch.n_fixed_steps
            f, _p = W.readme_hook(t, f, c["pars"])
            t, f = sch(t, f, c["dt"], c["pars"], hook=hook)
            snaps.append(f.uflat.copy())
            counts.append(sch.n_fixed_steps - n0)
        return "counts %s err %.2e sum %.10f" % (
            counts, rel_traj_err(np.array(snaps), g["readme_adaptive_ROS3PRw"]), snaps[-1].sum())
    attempt("readme adaptive ROS3PRw (Dirichlet)", lambda: adaptive(S.Dirichlet(U=(1, 0))))
    attempt("readme adaptive ROS3PRw (python hook)", lambda: adaptive(W.readme_hook))
    for acc in (1, 2):
        cb = W.burgers(2048, acc)
        attempt("burgers_up%d 2048 x50" % acc, lambda cb=cb, acc=acc: "%.2e" % rel_traj_err(
            run_fixed(gmodel(cb["model"]), S.ROS2(gmodel(cb["model"])), cb, 50, 10),
            g["burgers_up%d_2048" % acc]))
    for N in (2048, 1000):
        ck = W.kuramoto(N)
        attempt("ks %d x50" % N, lambda ck=ck, N=N: "%.2e" % rel_traj_err(
            run_fixed(gmodel("ks"), S.ROS3PRw(gmodel("ks"), **fx), ck, 50, 10), g["ks_%d" % N]))
    ck = W.kuramoto(512)
    attempt("ks 512 edge x20", lambda: "%.2e" % rel_traj_err(
        run_fixed(gmodel("ks"), S.ROS3PRw(gmodel("ks"), **fx), ck, 20, 5, pars=dict(periodic=False)),
        g["ks_512_edge"]))
    if not os.environ.get("TF_SKIP_FILM"):
        for theta in (1, .5):
            cf = W.film(1024, theta)
            attempt("film 1024 theta=%g x100" % theta, lambda cf=cf, theta=theta: "%.2e" % rel_traj_err(
                run_fixed(gmodel("film"), S.Theta(gmodel("film"), theta=theta), cf, 100, 20),
                g["film_1024_theta%g" % theta]))


if __name__ == "__main__":
    main()

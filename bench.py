#!/usr/bin/env python
"""Benchmark of the implicit method-of-lines hot path (BASELINE.json metric:
implicit grid-point*steps/s; F+J+solve fraction of the HBM roofline).

    python bench.py --gpus N --steps K --warmup W [--workload ensemble|ks|burgers|film]
    python bench.py --impl reference ...        # CPU arm: the reference's own numpy + SuperLU path

One "step" is one internal implicit step (one J build + factorisation + s stage
solves + update) of every system in the batch.  Inputs are synthetic
 and resident in HBM when the timed region starts; the timed
region is bracketed by a barrier + stream synchronisation and timed with CUDA
events on the launching stream, max over ranks.  Ensembles are sharded by member
with no data-path collective: the BASELINE config (32768 members in total) is split
contiguously over the ranks (strong scaling); the final gather of the states to rank 0
is timed separately (``final_gather``).
"""

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from triflow_b200 import workloads as W  # noqa: E402

# algorithmic bytes per node per internal step:
#   Rosenbrock s stages: Q = 8[(1+s)B + (1+4s+s(s-1)/2)v + (1+s)n_a],  B = v^2 w
#   (Theta is the 1-stage case of the same kernels: Q = 8[2B + 4v + n_a] in SURVEY;
#    the kernels move the Rosenbrock s=1 traffic, 8[2B + 5v], we quote SURVEY's.)
WORKLOADS = {
    "ensemble": dict(v=1, w=3, s=3, Q=224, scheme="ROS3PRw", cfg="configs[4]"),
    "burgers": dict(v=1, w=3, s=2, Q=152, scheme="ROS2", cfg="configs[1]"),
    "ks": dict(v=1, w=5, s=3, Q=288, scheme="ROS3PRw", cfg="configs[2]"),
    "film": dict(v=2, w=5, s=1, Q=384, scheme="Theta", cfg="configs[3]"),
}


def kernel_bytes(wk, family, stage=None):
    """Algorithmic bytes per node of one launch of a kernel family: the SURVEY
    §8d step model split over the launches (DESIGN.md 'Roofline accounting')."""
    v, w, s = wk["v"], wk["w"], wk["s"]
    B = v * v * w
    if family in ("sysstep", "gridstep"):        # whole step in one launch
        return wk["Q"]
    if family == "factor":
        return 8 * (B + v)                       # write factor, read U
    lshare = B * (w // 2) / w                    # L part of the factor
    if family == "fwd":                          # mean over stages
        return 8 * (lshare + 2 * v + v * (s - 1) / 2)
    if family == "bwd":
        return 8 * ((B - lshare) + 2 * v)
    return 0


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                 "--format=csv,noheader,nounits", "-lms", "20"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([time.monotonic()] + [c.strip() for c in line.split(",")])

    def wait_first(self, timeout=3.0):
        t0 = time.monotonic()
        while self.proc and not self.rows and time.monotonic() - t0 < timeout:
            time.sleep(0.01)

    def mark(self):
        """Start of the timed region (the sampler itself is started before the warm-up
        steps: nvidia-smi needs ~0.1 s to deliver its first sample)."""
        self.t_mark = time.monotonic()

    def __exit__(self, *a):
        if self.proc:
            time.sleep(0.15)
            self.proc.terminate()
            self.thread.join(timeout=2)

    def summary(self):
        sm, mx, reasons = [], None, set()
        t_mark = getattr(self, "t_mark", 0.0)
        timed = [r[1:] for r in self.rows if r[0] >= t_mark]
        window = "timed region"
        if len(timed) < 2:                       # very short region: same load since the warm-up
            timed, window = [r[1:] for r in self.rows], "warm-up + timed region"
        for r in timed:
            try:
                sm.append(float(r[0]))
                mx = float(r[1])
            except (ValueError, IndexError):
                continue
            for name, val in zip(["hw_slowdown", "hw_thermal_slowdown",
                                  "sw_thermal_slowdown", "sw_power_cap"], r[2:6]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx,
                "reasons": sorted(reasons), "samples": len(sm), "window": window}


def make_config(args, wk, mname, N, total_systems, dt, ws):
    """`config` of the JSON line: identical for the GPU arm and for --impl reference."""
    units = float(N) * total_systems
    return {"workload": "%s: %s" % (wk["cfg"], args.workload),
            "pde": mname, "scheme": wk["scheme"], "nodes": int(N),
            "systems": int(total_systems), "dt": dt, "fixed_step": True,
            "parallelism": ("members sharded x%d, no data-path collective, final gather" % ws
                            if args.workload == "ensemble" else "replicas x%d" % ws),
            "l2": "working set %.0f MB > 126 MB L2" % (units * wk["Q"] / 1e6)
            if units * wk["Q"] > 126e6 * ws else
            "working set %.0f MB fits L2; no flush (steps are dependent)" % (units * wk["Q"] / 1e6)}


def bind_near_gpu(index):
    """Pin this process to the CPU cores NVML reports as close to the GPU, BEFORE any pinned
    buffer is allocated (first touch then lands on the near NUMA node).  Returns a note."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cores = [64 * w + b for w, m in enumerate(words) for b in range(64) if (m >> b) & 1]
        cores = [c for c in cores if c < ncpu]
        if cores:
            os.sched_setaffinity(0, cores)
            return "bound to %d cores near GPU %d (%d..%d)" % (len(cores), index, cores[0], cores[-1])
    except Exception as e:  # noqa: BLE001
        return "not bound (%s)" % type(e).__name__
    return "not bound"


# ----------------------------------------------------------------- workloads
def build_problem(name, members, N=None):
    """Returns (model_name, scheme factory, x, fields, pars, hook, batch, N, dt)."""
    from triflow_b200 import schemes as S
    if name == "ensemble":
        c = W.ensemble(N or 4096, members)
        # one value per member, in the unambiguous (batch, 1) form: with 4096 members per GPU
        # (cfg 5 on 8 GPUs) a (batch,) array would have the length of the grid
        pars = {k: (np.asarray(v, dtype=float).reshape(-1, 1) if np.ndim(v) == 1 else v)
                for k, v in c["pars"].items()}
        return ("advdiff", lambda m: S.ROS3PRw(m, time_stepping=False), c["x"], c["fields"],
                pars, S.Dirichlet(U=(1.0, 0.0)), len(c["members"]), c["x"].size, c["dt"])
    if name == "ks":
        c = W.kuramoto(N or 2 ** 20)
        return ("ks", lambda m: S.ROS3PRw(m, time_stepping=False), c["x"], c["fields"],
                c["pars"], S.null_hook, 1, c["x"].size, c["dt"])
    if name == "burgers":
        c = W.burgers(N or 2 ** 17, 1)
        return ("burgers_up1", lambda m: S.ROS2(m), c["x"], c["fields"], c["pars"],
                S.null_hook, 1, c["x"].size, c["dt"])
    if name == "film":
        c = W.film(N or 2 ** 18)
        return ("film", lambda m: S.Theta(m, theta=1), c["x"], c["fields"], c["pars"],
                S.null_hook, 1, c["x"].size, c["dt"])
    raise SystemExit("unknown workload %r" % name)


def run_gpu(args):
    from triflow_b200 import _lib, distributed as D
    from triflow_b200.ensemble import Ensemble
    from triflow_b200.model import Model

    rank, ws = D.init()
    local = int(os.environ.get("LOCAL_RANK", "0"))
    affinity = bind_near_gpu(local)
    wk = WORKLOADS[args.workload]
    if args.workload == "ensemble":
        # BASELINE cfg 5: `--members` runs IN TOTAL, contiguous block per rank (strong scaling)
        lo, hi = D.shard(args.members, rank, ws)
        members = np.arange(lo, hi) % (W.ENSEMBLE_K * W.ENSEMBLE_C)
    else:
        members = None                          # replicas only
    mname, mk_scheme, x, fields, pars, hook, batch, N, dt = build_problem(
        args.workload, members, args.nodes)
    model = Model(**W.model_args(mname), compiler="cuda")
    scheme = mk_scheme(model)
    ens = Ensemble(model, scheme, x, fields, pars, hook=hook, batch=batch)
    lib, ctx = _lib.lib(), model._cuda.ctx
    units = float(N) * batch                    # nodes stepped per step on this rank

    with ClockSampler(local) as clocks:
        clocks.wait_first()
        # W untimed warm-up steps in ONE call, like the timed K: a run of >= 3 steps replays a
        # captured step graph, and building it (~0.2 ms) belongs to the warm-up, not to K steps
        if args.warmup > 0:
            ens.step(dt, args.warmup)
        ens.sync()
        # -- timed region: K steps, state resident in HBM
        launches0 = lib.tf_ctx_launch_count(ctx)
        D.barrier()
        ens.sync()
        clocks.mark()
        _lib.check(lib.tf_ctx_timer_start(ctx))
        import ctypes
        _lib.check(lib.tf_scheme_step(ens.state.h, scheme.handle, float(dt), args.steps, None))
        ms = ctypes.c_float()
        _lib.check(lib.tf_ctx_timer_stop(ctx, ctypes.byref(ms)))
        ens.sync()
    D.barrier()
    launches = lib.tf_ctx_launch_count(ctx) - launches0
    t_max = D.max_over_ranks(ms.value * 1e-3)
    total_units = D.sum_over_ranks(units)
    value = total_units * args.steps / t_max

    # -- per-kernel device times (separate short pass, CUDA events per launch)
    _lib.check(lib.tf_ctx_profile(ctx, 1))
    psteps = max(1, min(args.steps, 5))
    ens.step(dt, psteps)
    fam = {}
    for i, fname in enumerate(_lib.FAMILIES):
        fms, fn = ctypes.c_float(), ctypes.c_longlong()
        _lib.check(lib.tf_ctx_profile_read(ctx, i, ctypes.byref(fms), ctypes.byref(fn)))
        if fn.value:
            fam[fname] = (fms.value, fn.value)
    _lib.check(lib.tf_ctx_profile(ctx, 0))
    tot_ms = sum(v[0] for v in fam.values())
    dom = max((f for f in fam if kernel_bytes(wk, f) > 0), key=lambda f: fam[f][0])
    dom_ms, dom_n = fam[dom]
    peak, peak_src = peaks()
    achieved = kernel_bytes(wk, dom) * units / (dom_ms / dom_n * 1e-3) / 1e9
    step_gbs = wk["Q"] * units * args.steps / (ms.value * 1e-3) / 1e9
    # Per-launch figures that only a profiler can give (DRAM bytes, fp64 instruction counts)
    # are STATIC: taken from the committed ncu capture named in `static_source`, scaled by the
    # nodes of this run -- not measured in this process.
    static, spath = {}, os.path.join(ROOT, "profiles", "r2b_static.json")
    if os.path.exists(spath):
        with open(spath) as f:
            static = json.load(f)
    sk = static.get(args.workload, {}).get(dom, {})
    traffic = sk.get("dram_bytes_per_node") and sk["dram_bytes_per_node"] * units
    roofline = {"bound": sk.get("bound", "hbm"), "kernel": "tf_k_" + dom, "achieved": round(achieved, 1),
                "peak": peak, "unit": "GB/s", "frac": round(achieved / peak, 4),
                "traffic": traffic,
                "traffic_source": ("static ncu capture: " + sk["source"]) if traffic else None,
                "peak_source": peak_src,
                "limiter": sk.get("limiter"),
                "kernel_share_of_step": round(dom_ms / tot_ms, 3),
                "step_achieved": round(step_gbs, 1), "step_frac": round(step_gbs / peak, 4),
                "bytes_per_node_step": wk["Q"],
                "family_ms_per_step": {k: round(v[0] / psteps, 4) for k, v in fam.items()}}
    # fp64 pipe: measured peak of this device (DFMA / s, a probe kernel in the library) against
    # the fp64 operations the kernel executes per node and step (static, same capture)
    dfma_peak = ctypes.c_double()
    _lib.check(lib.tf_ctx_fp64_peak(ctx, ctypes.byref(dfma_peak)))
    roofline["fp64_peak_dfma_per_s"] = dfma_peak.value
    if sk.get("fp64_inst_per_node_step"):
        rate = sk["fp64_inst_per_node_step"] * units / (dom_ms / dom_n * 1e-3)
        roofline["fp64"] = {"achieved_inst_per_s": rate, "peak_inst_per_s": dfma_peak.value,
                            "frac": round(rate / dfma_peak.value, 4),
                            "inst_per_node_step": sk["fp64_inst_per_node_step"],
                            "source": "static ncu capture: " + sk["source"]}
    if roofline["bound"] != "hbm":
        roofline["note"] = ("achieved / frac quote the ALGORITHMIC bytes of SURVEY 8d (Q) over the "
                            "kernel time against the HBM peak, as BASELINE.json asks; this kernel "
                            "keeps the factor and the stage vectors on the SM, its DRAM traffic is "
                            "`traffic`, and what bounds it is the fp64 pipe / issue rate (`fp64`)")

    # -- final gather of the states to rank 0 (the only inter-GPU traffic of an ensemble)
    h_final = _lib.pinned_empty((batch, N * model._nvar))
    if args.workload == "ensemble" and ws > 1:
        # the first collective of a process group builds the NCCL communicator (hundreds of
        # ms): not part of the gather
        import torch
        D.gather_members(torch.zeros((batch, 1), dtype=torch.float64, device="cuda"), args.members)
        h_full = _lib.pinned_empty((args.members, N * model._nvar)) if rank == 0 else None
    D.barrier()
    t0 = time.perf_counter()
    if args.workload == "ensemble" and ws > 1:
        # sharded ensemble: unpack on the device, NCCL gather over NVLink, one copy to the host
        gathered = D.gather_members(ens.download_to_torch(), args.members, out=h_full)
    else:
        gathered = ens.download(out=h_final)
    t_gather = D.max_over_ranks(time.perf_counter() - t0)
    final_gather = {"ms": t_gather * 1e3, "bytes": int(8 * total_units * model._nvar),
                    "what": ("unpack on the device, NCCL gather to rank 0, one device -> pinned host copy there"
                             if ws > 1 else "device -> host download"),
                    "value_incl_gather": total_units * args.steps / (t_max + t_gather)}
    if rank == 0 and args.workload == "ensemble":
        assert gathered.shape[0] == args.members
        if ws > 1:                                # this rank's block arrived where it belongs
            assert np.array_equal(gathered[lo:hi], ens.download(out=h_final)), "gather misplaced a block"
    del gathered

    # -- end to end through the public API with HOST buffers: every step uploads the
    #    unknowns from pinned memory, steps, and downloads the result (the reference's
    #    scheme(t, fields, dt, pars) convention); member blocks are pipelined over streams
    from triflow_b200.ensemble import HostPipeline
    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    nv = model._nvar
    h_in = _lib.pinned_empty((batch, N * nv))
    h_out = _lib.pinned_empty((batch, N * nv))
    h_in[:] = ens.download()
    u_before = h_in[0].copy()
    pipe = HostPipeline(model, scheme, x, fields, pars, hook=hook, batch=batch,
                        groups=args.e2e_groups)
    pipe.step_host(h_in, h_out, dt)             # untimed warm-up (result discarded)
    D.barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        pipe.step_host(h_in, h_out, dt)
        h_in, h_out = h_out, h_in
    t_rank = time.perf_counter() - t0
    t_e2e = D.max_over_ranks(t_rank)
    link = 8.0 * N * nv * batch * e2e_steps / t_rank / 1e9     # GB/s each way on this rank
    e2e = {"value": total_units * e2e_steps / t_e2e, "unit": "grid-point*steps/s",
           "h2d_bytes_per_step": int(8 * N * nv * batch), "d2h_bytes_per_step": int(8 * N * nv * batch),
           "steps": e2e_steps, "ms_per_step": t_e2e * 1e3 / e2e_steps,
           "api": "HostPipeline.step_host (%d member blocks on their own streams)" % len(pipe.parts),
           "bytes_are": "per rank",
           "link_gbs_each_way": {"rank0": round(link, 2),
                                 "min_over_ranks": round(-D.max_over_ranks(-link), 2),
                                 "sum_over_ranks": round(D.sum_over_ranks(link), 2)},
           "host": affinity,
           "limiter": "PCIe / host memory: upload and download of every step run concurrently"}
    assert not np.array_equal(u_before, h_in[0]), "the end-to-end pass did not advance the state"
    pipe.close()
    status = ens.state.status()
    assert not status.any(), "factorisation failed in the bench"
    assert np.isfinite(h_in).all(), "non-finite state after the bench"

    extras = {}
    if args.workload == "ensemble" and ens.state.variant.lowered.jacobian_is_constant:
        # same workload with the factorisation kept across steps (legitimate for this
        # linear model; NOT the headline: the reference refactorises)
        _lib.check(lib.tf_state_set_factor_reuse(ens.state.h, 1))
        ens.step(dt, 2)
        ens.sync()
        _lib.check(lib.tf_ctx_timer_start(ctx))
        ens.step(dt, args.steps)
        ms2 = ctypes.c_float()
        _lib.check(lib.tf_ctx_timer_stop(ctx, ctypes.byref(ms2)))
        _lib.check(lib.tf_state_set_factor_reuse(ens.state.h, 0))
        t2 = D.max_over_ranks(ms2.value * 1e-3)
        extras["factor_reuse"] = {"value": total_units * args.steps / t2,
                                  "ms_per_step": t2 * 1e3 / args.steps,
                                  "note": "factor kept while gamma*dt is constant"}
    ens.state.close()
    if rank == 0 and ws == 1 and args.others:
        for other in ("ks", "burgers", "film"):
            if other == args.workload:
                continue
            try:
                extras[other] = quick_measure(other, min(args.steps, 10))
            except Exception as e:  # noqa: BLE001  (never lose the headline line)
                extras[other] = {"error": "%s: %s" % (type(e).__name__, str(e)[:200])}
    if ws > 1 and args.others:
        try:                                    # every rank takes part
            extras["ks_one_grid"] = slab_measure(ws, min(args.steps, 10))
        except Exception as e:  # noqa: BLE001
            extras["ks_one_grid"] = {"error": "%s: %s" % (type(e).__name__, str(e)[:200])}
    cpu = None
    if rank == 0 and ws == 1 and not args.no_cpu:
        try:
            os.sched_setaffinity(0, range(os.cpu_count() or 1))   # the CPU arm uses every core
        except OSError:
            pass
        cpu = cpu_baseline(args.workload, args.cpu_seconds)
    if rank == 0:
        out = {
            "metric": "implicit grid-point*steps/s", "value": value,
            "unit": "grid-point*steps/s", "n_gpus": ws, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": t_max * 1e3 / args.steps,
            "higher_is_better": True,
            "scaling": "strong" if args.workload == "ensemble" else "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": make_config(args, wk, mname, N,
                                  args.members if args.workload == "ensemble" else ws, dt, ws),
            "final_gather": final_gather,
            "clocks": clocks.summary(), "e2e": e2e, "gpu_launches": int(launches),
            "roofline": roofline, "cpu_baseline": cpu, "other_workloads": extras,
        }
        _emit(out)
    D.finalize()


def slab_measure(ws, steps):
    """ONE Kuramoto-Sivashinsky grid of ws x 2^20 nodes stepped by all GPUs together (SURVEY K7:
    slab per GPU, halo / scan / border words read from the peer GPU over NVLink inside the step
    kernel).  Every rank calls this; device-timed with CUDA events, max over ranks."""
    import ctypes
    from triflow_b200 import _lib, distributed as D
    from triflow_b200.model import Model
    wk = WORKLOADS["ks"]
    N = ws << 20
    c = W.kuramoto(N)
    model = Model(**W.model_args("ks"), compiler="cuda")
    from triflow_b200 import schemes as S
    scheme = S.ROS3PRw(model, time_stepping=False)
    g = D.SlabGrid(model, scheme, c["x"], c["fields"], c["pars"])
    lib, ctx = _lib.lib(), model._cuda.ctx
    g.step(c["dt"], 3)
    g.sync()
    D.barrier()
    _lib.check(lib.tf_ctx_timer_start(ctx))
    g.step(c["dt"], steps)
    ms = ctypes.c_float()
    _lib.check(lib.tf_ctx_timer_stop(ctx, ctypes.byref(ms)))
    g.sync()
    t = D.max_over_ranks(ms.value * 1e-3)
    local = g.download()
    ok = bool(np.isfinite(local).all())
    tiles = (g.states[0].tiles_local, g.states[0].tiles_total)
    g.close()
    peak, _ = peaks()
    rate = float(N) * steps / t
    return {"config": "configs[2] x %d: one grid over %d GPUs" % (ws, ws), "pde": "ks", "scheme": wk["scheme"],
            "nodes": N, "value": rate, "ms_per_step": t * 1e3 / steps, "finite": ok,
            "tiles_per_gpu": tiles[0], "tiles": tiles[1],
            "exchange": "tagged 16-byte words read from the peer GPU's record area over NVLink inside "
                        "the step kernel (no collective)",
            "bytes_per_node_step": wk["Q"], "step_frac_of_all_gpus": round(wk["Q"] * rate / 1e9 / (peak * ws), 4)}


def quick_measure(workload, steps):
    """Short device-resident measurement of another BASELINE workload (1 GPU)."""
    import ctypes
    from triflow_b200 import _lib
    from triflow_b200.ensemble import Ensemble
    from triflow_b200.model import Model
    wk = WORKLOADS[workload]
    mname, mk_scheme, x, fields, pars, hook, batch, N, dt = build_problem(workload, None)
    model = Model(**W.model_args(mname), compiler="cuda")
    scheme = mk_scheme(model)
    ens = Ensemble(model, scheme, x, fields, pars, hook=hook, batch=batch)
    lib, ctx = _lib.lib(), model._cuda.ctx
    ens.step(dt, 3)
    ens.sync()
    _lib.check(lib.tf_ctx_timer_start(ctx))
    _lib.check(lib.tf_scheme_step(ens.state.h, scheme.handle, float(dt), steps, None))
    ms = ctypes.c_float()
    _lib.check(lib.tf_ctx_timer_stop(ctx, ctypes.byref(ms)))
    peak, _ = peaks()
    rate = float(N) * batch * steps / (ms.value * 1e-3)
    assert np.isfinite(ens.download()).all()
    ens.state.close()
    return {"config": wk["cfg"], "pde": mname, "scheme": wk["scheme"], "nodes": N,
            "value": rate, "ms_per_step": ms.value / steps,
            "bytes_per_node_step": wk["Q"], "step_frac": round(wk["Q"] * rate / 1e9 / peak, 4)}


# -------------------------------------------------------------- CPU baseline
def _cpu_kit():
    """The reference's own CPU path -- its unmodified core files from baseline/_ref (copied by
    build(); kind "reference") -- or, if they are not there, the oracle port (kind "port").
    Returns (Model class, compiler argument, schemes module, kind)."""
    try:
        from oracle import ref_loader
        if ref_loader.reference_root() is not None:
            mods = ref_loader.load_reference()
            return mods["model"].Model, "numpy", mods["schemes"], "reference"
    except Exception:  # noqa: BLE001
        pass
    from oracle import schemes as O
    from oracle.numpy_compiler import numpy_compiler
    from triflow_b200.model import Model
    return Model, numpy_compiler, O, "port"


_CPU = {}


def _cpu_init():
    """Per worker process, outside the timed region: build the model once (SymPy front-end +
    lambdify take seconds; pickling a reference model drops its compiler, model.py:579-583)."""
    Model, comp, schemes, kind = _cpu_kit()
    _CPU.update(model=Model(**W.model_args("advdiff"), compiler=comp), schemes=schemes, kind=kind)


def _cpu_member(job):
    r, N, steps = job
    m, schemes = _CPU["model"], _CPU["schemes"]
    c = W.ensemble(N, [r])
    pars = dict(k=float(c["pars"]["k"][0]), c=float(c["pars"]["c"][0]), periodic=False)
    f = m.fields_template(x=c["x"], **c["fields"])
    sch = schemes.ROS3PRw(m, time_stepping=False)
    t = 0.0
    for _ in range(steps):
        t, f = sch(t, f, c["dt"], pars, hook=W.readme_hook)
    return _CPU["kind"]


def cpu_rate(workload, budget_s, cores):
    """The reference CPU path (numpy compiler + SciPy SuperLU) on the same workload, on host
    cores.  Only stepping is timed: model construction and process start-up are outside."""
    if workload == "ensemble":
        import multiprocessing as mp
        N, steps = 4096, 20
        with mp.get_context("fork").Pool(cores, initializer=_cpu_init) as pool:
            pool.map(_cpu_member, [(0, N, 1)] * cores)               # workers up, models built
            t0 = time.perf_counter()
            kind = pool.map(_cpu_member, [(0, N, 2)] * cores)[0]
            per = (time.perf_counter() - t0) / 2
            nmem = max(cores, int(budget_s / max(per * steps, 1e-3)) * cores)
            nmem = min(nmem, 512 * cores)
            jobs = [(int(r), N, steps)
                    for r in np.linspace(0, W.ENSEMBLE_K * W.ENSEMBLE_C - 1, nmem)]
            t0 = time.perf_counter()
            pool.map(_cpu_member, jobs, chunksize=max(1, nmem // (4 * cores)))
            wall = time.perf_counter() - t0
        return (N * nmem * steps / wall, kind,
                "%d of 32768 members x %d steps over %d processes (stepping only)" % (nmem, steps, cores))
    Model, comp, schemes, kind = _cpu_kit()
    mk = {"ks": (W.kuramoto, "ks", lambda m: schemes.ROS3PRw(m, time_stepping=False)),
          "burgers": (lambda N=None: W.burgers(N or 2 ** 17, 1), "burgers_up1", schemes.ROS2),
          "film": (W.film, "film", lambda m: schemes.Theta(m, theta=1))}[workload]
    c = mk[0]()
    m = Model(**W.model_args(mk[1]), compiler=comp)
    f = m.fields_template(x=c["x"], **c["fields"])
    sch = mk[2](m)
    t, n, t0 = 0.0, 0, time.perf_counter()
    while True:
        t, f = sch(t, f, c["dt"], c["pars"])
        n += 1
        if time.perf_counter() - t0 > budget_s or n >= 20:
            break
    wall = time.perf_counter() - t0
    return (c["x"].size * n / wall, kind, "%d steps at full size N=%d, 1 process "
            "(the path is single-threaded)" % (n, c["x"].size))


def cpu_baseline(workload, budget_s):
    cores = (os.cpu_count() or 1) if workload == "ensemble" else 1
    v, kind, sample = cpu_rate(workload, budget_s, cores)
    return {"value": v, "unit": "grid-point*steps/s", "cores": cores, "kind": kind,
            "sample": sample + "; rate extrapolated linearly"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    ws = int(os.environ.get("WORLD_SIZE", "1"))
    wk = WORKLOADS[args.workload]
    dims = {"ensemble": ("advdiff", args.nodes or 4096, args.members, 0.025),
            "ks": ("ks", args.nodes or 2 ** 20, ws, 0.2),
            "burgers": ("burgers_up1", args.nodes or 2 ** 17, ws, 0.1),
            "film": ("film", args.nodes or 2 ** 18, ws, 0.05)}[args.workload]
    cores = (os.cpu_count() or 1) if args.workload == "ensemble" else 1
    budget = max(2.0, min(25.0, 120.0 / max(1, args.steps + args.warmup)))
    for _ in range(min(args.warmup, 1)):
        cpu_rate(args.workload, 1.0, cores)
    vals, sample, kind = [], "", "port"
    for _ in range(max(1, min(args.steps, 3))):
        v, kind, sample = cpu_rate(args.workload, budget, cores)
        vals.append(v)
    v = float(np.mean(vals))
    _emit({
        "impl": "reference", "metric": "implicit grid-point*steps/s", "value": v,
        "unit": "grid-point*steps/s", "n_gpus": ws,
        "steps": args.steps, "warmup": args.warmup, "higher_is_better": True,
        "scaling": "strong" if args.workload == "ensemble" else "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": make_config(args, wk, dims[0], dims[1], dims[2], dims[3], ws),
        "cpu_baseline": {"value": v, "unit": "grid-point*steps/s", "cores": cores,
                         "kind": kind, "sample": sample},
        "e2e": {"value": v, "unit": "grid-point*steps/s", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0}})


_emit = None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="ensemble", choices=sorted(WORKLOADS))
    ap.add_argument("--members", type=int, default=W.ENSEMBLE_K * W.ENSEMBLE_C,
                    help="ensemble members IN TOTAL (sharded over the ranks)")
    ap.add_argument("--nodes", type=int, default=None)
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--e2e-groups", type=int, default=16,
                    help="member blocks of the pipelined host path")
    ap.add_argument("--cpu-seconds", type=float, default=15.0)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-others", dest="others", action="store_false",
                    help="skip the short ks/burgers/film measurements added to the line")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    # stdout carries exactly one JSON line: anything a library prints meanwhile (NCCL's
    # version banner, ...) is routed to stderr, and stdout is restored for the result
    sys.stdout.flush()
    saved = os.dup(1)
    os.dup2(2, 1)
    global _emit

    def _emit(obj):
        sys.stdout.flush()
        os.dup2(saved, 1)
        print(json.dumps(obj), flush=True)
        os.dup2(2, 1)

    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)
    sys.stdout.flush()
    os.dup2(saved, 1)


if __name__ == "__main__":
    main()

"""ctypes binding of ``libtriflow_b200.so`` (``include/triflow_b200.h``) and the
nvcc build of the library and of the per-model cubins.

There is no CPU fallback anywhere in this package: if the library is missing it
is built with nvcc; if no CUDA device is present, ``tf_ctx_create`` fails and
:class:`CudaUnavailable` is raised.
"""

import ctypes as C
import hashlib
import os
import shutil
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_PATH = os.path.join(HERE, "libtriflow_b200.so")
CACHE = os.path.join(HERE, "_kcache")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]

TF_OK, TF_EINVAL, TF_ECUDA, TF_EMAXITER, TF_EDTMIN, TF_ESINGULAR = range(6)
FAMILIES = ["factor", "border_fill", "fwd", "border_solve", "bwd", "update", "hook",
            "pack", "eval", "sysstep", "gridstep"]


class CudaUnavailable(RuntimeError):
    pass


class ModelDesc(C.Structure):
    _fields_ = [(n, C.c_int) for n in (
        "nvar", "nhelp", "half_width", "nnz", "n_const", "n_nodepar", "uses_x",
        "chunk_nodes", "warps")]


def _nvcc():
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise CudaUnavailable("nvcc not found: cannot build triflow_b200 kernels")
    return exe


def _sources_digest(paths):
    h = hashlib.sha1()
    for p in paths:
        with open(p, "rb") as f:
            h.update(f.read())
    return h.hexdigest()[:16]


def build_library(force=False):
    """Compile csrc/tf_host.cu -> libtriflow_b200.so (static cudart, sm_100a)."""
    srcs = [os.path.join(CSRC, "tf_host.cu"), os.path.join(CSRC, "tf_params.h"),
            os.path.join(os.path.dirname(HERE), "include", "triflow_b200.h")]
    stamp = LIB_PATH + ".stamp"
    flags = ["-O2", "-std=c++17", "-shared", "-Xcompiler", "-fPIC", "-cudart", "static", *ARCH,
             "-lineinfo", "-diag-suppress", "177"]
    digest = _sources_digest(srcs)
    if (not force and os.path.exists(LIB_PATH) and os.path.exists(stamp)
            and open(stamp).read().split("|")[0] == digest):
        return LIB_PATH
    try:
        ver = subprocess.run([_nvcc(), "--version"], capture_output=True, text=True).stdout.split()[-1]
    except Exception:  # noqa: BLE001
        ver = "?"
    # several ranks may build at once on a fresh checkout: private temporary, atomic rename
    tmp = "%s.tmp%d" % (LIB_PATH, os.getpid())
    subprocess.check_call([_nvcc(), *flags, "-o", tmp, srcs[0]])
    os.replace(tmp, LIB_PATH)
    with open(stamp + ".tmp%d" % os.getpid(), "w") as f:
        f.write("%s|%s|%s" % (digest, " ".join(flags), ver))
    os.replace(stamp + ".tmp%d" % os.getpid(), stamp)
    return LIB_PATH


_KERNEL_SRCS = ["tf_kernels.cuh", "tf_sysstep.cuh", "tf_gridstep.cuh", "tf_band.h", "tf_params.h", "tf_model_prelude.h"]


def build_cubin(header, chunk_nodes, warps, fast_div=False):
    """Compile a generated model header + tf_kernels.cuh to an sm_100a cubin."""
    os.makedirs(CACHE, exist_ok=True)
    digest = _sources_digest([os.path.join(CSRC, s) for s in _KERNEL_SRCS])
    minb = int(os.environ.get("TF_MINB", "0"))          # 0: the kernels' own default
    extra = os.environ.get("TF_CFLAGS", "").split()          # tuning knobs (-DTF_...)
    key = hashlib.sha1(("%s|%s|%d|%d|%d|%s" % (header, digest, chunk_nodes, minb,
                                               int(fast_div), extra)).encode()).hexdigest()[:20]
    cubin = os.path.join(CACHE, "m_%s.cubin" % key)
    if os.path.exists(cubin):
        return cubin
    src = os.path.join(CACHE, "m_%s.cu" % key)
    with open(src, "w") as f:
        f.write(header)
        f.write('#include "tf_kernels.cuh"\n')
    tmp = cubin + ".tmp%d" % os.getpid()
    cmd = [_nvcc(), *ARCH, "-O3", "-std=c++17", "-lineinfo", "-diag-suppress", "550", "-I", CSRC,
           "-DTF_M=%d" % chunk_nodes, "-DTF_FAST_DIV=%d" % int(fast_div),
           *(["-DTF_MINB=%d" % minb] if minb else []), *extra, "-cubin", "-o", tmp, src]
    subprocess.check_call(cmd)
    os.replace(tmp, cubin)
    return cubin


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    path = build_library()
    L = C.CDLL(path)
    vp, i, d, dp = C.c_void_p, C.c_int, C.c_double, C.POINTER(C.c_double)
    L.tf_last_error.restype = C.c_char_p
    sig = {
        "tf_ctx_create": [i, C.POINTER(vp)], "tf_ctx_destroy": [vp], "tf_ctx_sync": [vp],
        "tf_ctx_set_async": [vp, i],
        "tf_host_alloc": [C.c_size_t, C.POINTER(vp)], "tf_host_free": [vp],
        "tf_host_alloc_wc": [C.c_size_t, C.POINTER(vp)],
        "tf_ctx_fp64_peak": [vp, dp],
        "tf_ring_create": [vp, i, C.POINTER(vp)], "tf_ring_destroy": [vp],
        "tf_ring_push": [vp, d], "tf_ring_pop": [vp, i, C.POINTER(dp), dp],
        "tf_ring_release": [vp],
        "tf_model_load": [vp, vp, C.c_size_t, C.POINTER(ModelDesc), C.POINTER(vp)],
        "tf_model_unload": [vp],
        "tf_model_read_symbol": [vp, C.c_char_p, vp, C.c_size_t],
        "tf_state_create": [vp, vp, i, i, i, C.POINTER(vp)], "tf_state_destroy": [vp],
        "tf_state_upload": [vp, dp, dp, dp, dp, dp], "tf_state_download": [vp, dp],
        "tf_state_download_device": [vp, vp],
        "tf_eval_F": [vp, dp], "tf_eval_J": [vp, dp],
        "tf_scheme_create": [vp, i, dp, dp, dp, dp, C.POINTER(vp)],
        "tf_scheme_destroy": [vp],
        "tf_hook_set_dirichlet": [vp, i, i, d, i, d], "tf_hook_clear": [vp],
        "tf_scheme_step": [vp, vp, d, i, dp],
        "tf_scheme_advance": [vp, vp, d, d, d, d, i, d, i, dp, C.POINTER(i), dp],
        "tf_state_status": [vp, C.POINTER(i)], "tf_state_set_factor_reuse": [vp, i],
        "tf_state_set_fusion": [vp, i],
        "tf_state_create_slab": [vp, vp, i, i, i, i, C.POINTER(vp)],
        "tf_state_slab_info": [vp, C.POINTER(i), C.POINTER(i), C.POINTER(i), C.POINTER(i)],
        "tf_state_slab_export": [vp, vp], "tf_state_slab_attach": [vp, vp],
        "tf_state_slab_attach_local": [vp, C.POINTER(vp)],
        "tf_ensemble_advance": [vp, vp, d, d, d, d, i, d, dp, C.POINTER(i), C.POINTER(i)],
        "tf_ensemble_richardson": [vp, vp, i, d, d, d, i, d, d, d, i, d, dp, dp, C.POINTER(i),
                                   C.POINTER(i), C.POINTER(i)],
        "tf_ctx_timer_start": [vp], "tf_ctx_timer_stop": [vp, C.POINTER(C.c_float)],
        "tf_ctx_profile": [vp, i],
        "tf_ctx_profile_read": [vp, i, C.POINTER(C.c_float), C.POINTER(C.c_longlong)],
    }
    for name, args in sig.items():
        fn = getattr(L, name)
        fn.argtypes = args
        fn.restype = i
    L.tf_ctx_launch_count.argtypes = [vp]
    L.tf_ctx_launch_count.restype = C.c_longlong
    _lib = L
    return L


EXPORTS = ["tf_last_error", "tf_ctx_create", "tf_ctx_destroy", "tf_ctx_sync", "tf_ctx_set_async",
           "tf_host_alloc", "tf_host_alloc_wc", "tf_ctx_fp64_peak", "tf_ring_create",
           "tf_ring_destroy", "tf_ring_push", "tf_ring_pop", "tf_ring_release",
           "tf_host_free", "tf_model_load", "tf_model_unload", "tf_model_read_symbol", "tf_state_create",
           "tf_state_destroy", "tf_state_upload", "tf_state_download", "tf_state_download_device", "tf_eval_F",
           "tf_eval_J", "tf_scheme_create", "tf_scheme_destroy", "tf_hook_set_dirichlet",
           "tf_hook_clear", "tf_scheme_step", "tf_scheme_advance", "tf_state_status",
           "tf_state_set_factor_reuse", "tf_state_set_fusion", "tf_ensemble_advance",
           "tf_state_create_slab", "tf_state_slab_info", "tf_state_slab_export",
           "tf_state_slab_attach", "tf_state_slab_attach_local",
           "tf_ensemble_richardson",
           "tf_ctx_launch_count", "tf_ctx_timer_start", "tf_ctx_timer_stop",
           "tf_ctx_profile", "tf_ctx_profile_read"]


def check(rc):
    """Map a status code to the reference's exception types."""
    if rc == TF_OK:
        return
    msg = lib().tf_last_error().decode()
    if rc in (TF_EMAXITER, TF_EDTMIN):
        raise RuntimeError(msg)           # reference schemes.py:229-238
    if rc == TF_ESINGULAR:
        raise RuntimeError(msg)           # SciPy raises RuntimeError on singular A
    if rc == TF_EINVAL:
        raise ValueError(msg)
    raise CudaUnavailable(msg)


def dptr(a):
    if a is None:
        return None
    assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.POINTER(C.c_double))


_ctx = {}


def context(device=None):
    """Process-wide context per device (default: LOCAL_RANK or 0)."""
    if device is None:
        device = int(os.environ.get("LOCAL_RANK", "0"))
    if device not in _ctx:
        h = C.c_void_p()
        check(lib().tf_ctx_create(device, C.byref(h)))
        _ctx[device] = h
    return _ctx[device]


def new_context(device=None):
    """An additional context (own stream) on a device; the caller keeps the handle."""
    if device is None:
        device = int(os.environ.get("LOCAL_RANK", "0"))
    h = C.c_void_p()
    check(lib().tf_ctx_create(device, C.byref(h)))
    return h


_pinned = []


def pinned_empty(shape, dtype=np.float64, write_combined=False):
    """numpy array over page-locked host memory (for upload / download buffers).
    ``write_combined``: for buffers the host only writes (upload sources)."""
    n = int(np.prod(shape)) * np.dtype(dtype).itemsize
    p = C.c_void_p()
    check((lib().tf_host_alloc_wc if write_combined else lib().tf_host_alloc)(n, C.byref(p)))
    buf = (C.c_char * max(n, 1)).from_address(p.value)
    arr = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)
    _pinned.append((p, buf))  # kept for the life of the process
    return arr

"""ctypes binding of the C ABI in ``include/ude_b200.h`` (+ the in-tree nvcc build).

The shared library ``csrc/libude_b200.so`` is plain CUDA C++ (no torch types in
any signature).  This module is the *only* place Python touches it.  There is
no fallback: if the library is missing or no B200 is visible the calls raise.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
SO_PATH = os.environ.get("UDE_B200_LIB") or os.path.join(CSRC, "libude_b200.so")   # env override: kernel experiments
INCLUDE = os.path.join(os.path.dirname(_HERE), "include")

UDE_MAX_LAYERS = 8
UDE_OK = 0
UDE_ERR_CAPACITY = -7
FLAG_WEIGHTS_PACKED = 16
FLAG_BAYES = 32
KIND = {"Fp": 0, "Fa": 1, "FaFp": 2}
METHOD = {"rk4": 0, "dopri5": 1}


class UdeDesc(C.Structure):
    _fields_ = [
        ("kind", C.c_int32),
        ("n_regions", C.c_int32),
        ("latent_dim", C.c_int32),
        ("n_rate_layers", C.c_int32),
        ("rate_sizes", C.c_int32 * UDE_MAX_LAYERS),
        ("n_aug_layers", C.c_int32),
        ("aug_sizes", C.c_int32 * UDE_MAX_LAYERS),
        ("fa_w", C.c_float),
        ("n_groups", C.c_int32),
        ("flags", C.c_int32),
        ("batch", C.c_int64),
    ]


class UdePlan(C.Structure):
    _fields_ = [
        ("n_theta", C.c_int64),
        ("workspace_bytes", C.c_int64),
        ("tape_bytes", C.c_int64),
        ("tile_traj", C.c_int32),
        ("n_ctas", C.c_int32),
        ("smem_bytes_fwd", C.c_int32),
        ("smem_bytes_bwd", C.c_int32),
        ("weights_in_smem", C.c_int32),
        ("path", C.c_int32),
    ]


class UdeReparam(C.Structure):
    _fields_ = [
        ("eps", C.c_void_p),
        ("mean", C.c_void_p),
        ("std", C.c_void_p),
        ("n_samples", C.c_int32),
        ("batch", C.c_int32),
    ]


class UdeDecoder(C.Structure):
    _fields_ = [
        ("weight", C.c_void_p),
        ("bias", C.c_void_p),
        ("n_out", C.c_int32),
    ]


class UdeDopri5Opts(C.Structure):
    _fields_ = [
        ("rtol", C.c_double),
        ("atol", C.c_double),
        ("first_step", C.c_double),
        ("safety", C.c_double),
        ("ifactor", C.c_double),
        ("dfactor", C.c_double),
        ("max_steps", C.c_int32),
        ("max_attempts", C.c_int32),
        ("shared_step", C.c_int32),
        ("reserved", C.c_int32),
    ]


# translation units (compiled in parallel, linked into one .so) and what each one includes
_UNIT_DEPS = {
    "ude_host.cu": ["ude_engine.cuh", "ude_rk4.cuh", "ude_dopri5.cuh", "ude_dopri5_host.inl", "ude_model.h",
                    "ude_host_common.h", "ude_tc_model.h"],
    "ude_tc_host.cu": ["ude_engine.cuh", "ude_model.h", "ude_host_common.h", "ude_tc.cuh", "ude_tc_model.h"],
}
_UNITS = list(_UNIT_DEPS)
_SOURCES = _UNITS + sorted({d for deps in _UNIT_DEPS.values() for d in deps})
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-split-compile", "0",
    "-Xcompiler", "-fPIC,-ffp-contract=off",
]


def _stale() -> bool:
    if not os.path.exists(SO_PATH):
        return True
    so_m = os.path.getmtime(SO_PATH)
    deps = [os.path.join(CSRC, s) for s in _SOURCES] + [os.path.join(INCLUDE, "ude_b200.h")]
    return any(os.path.exists(d) and os.path.getmtime(d) > so_m for d in deps)


def build(force: bool = False, verbose: bool = False, extra_flags=()) -> str:
    """Compile csrc/ for sm_100a with nvcc into csrc/libude_b200.so (in-tree)."""
    if not force and not _stale():
        return SO_PATH
    nvcc = os.environ.get("NVCC", "nvcc")
    objdir = os.path.join(os.path.dirname(_HERE), "build")
    os.makedirs(objdir, exist_ok=True)
    procs, objs = [], []
    for unit in _UNITS:
        obj = os.path.join(objdir, unit.replace(".cu", ".o"))
        deps = [os.path.join(CSRC, d) for d in [unit, *_UNIT_DEPS[unit]]] + [os.path.join(INCLUDE, "ude_b200.h")]
        if not force and os.path.exists(obj) and all(os.path.getmtime(d) <= os.path.getmtime(obj) for d in deps):
            objs.append(obj)        # object still current (objects live in build/, which is not tracked)
            continue
        cmd = [nvcc, *NVCC_FLAGS, *extra_flags, "-c", "-o", obj, os.path.join(CSRC, unit)]
        if verbose:
            cmd[1:1] = ["-Xptxas", "-v"]
            print(" ".join(cmd))
        procs.append((obj, subprocess.Popen(cmd, cwd=CSRC, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for obj, pr in procs:
        out, _ = pr.communicate()
        if pr.returncode != 0:
            raise RuntimeError("nvcc failed:\n" + out)
        if verbose:
            print(out)
        objs.append(obj)
    res = subprocess.run([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", SO_PATH, *objs],
                         cwd=CSRC, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc link failed:\n" + res.stdout + res.stderr)
    return SO_PATH


_lib = None
_lock = threading.Lock()

EXPORTS = [
    "ude_abi_version", "ude_strerror", "ude_last_cuda_error", "ude_theta_size", "ude_macs_per_eval", "ude_plan",
    "ude_rk4_grid", "ude_rk4_forward", "ude_rk4_backward", "ude_dopri5_forward", "ude_dopri5_backward",
    "ude_measure_ffma_tflops", "ude_rk4_path", "ude_rk4_forward_reparam", "ude_rk4_backward_reparam",
    "ude_rk4_forward_decoded", "ude_rk4_backward_decoded", "ude_rk4_forward_bayes", "ude_rk4_backward_bayes",
    "ude_rk4_forward_bayes_decoded", "ude_rk4_backward_bayes_decoded",
]


def lib() -> C.CDLL:
    """Load the C-ABI library (never builds implicitly on a GPU box; never falls back)."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(SO_PATH):
            raise RuntimeError(
                f"{SO_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a).  There is no CPU or PyTorch fallback for the fused UDE solve.")
        L = C.CDLL(SO_PATH)
        vp, i32, i64, f32 = C.c_void_p, C.c_int32, C.c_int64, C.c_float
        L.ude_abi_version.restype = C.c_int
        L.ude_strerror.restype = C.c_char_p
        L.ude_strerror.argtypes = [C.c_int]
        L.ude_last_cuda_error.restype = C.c_char_p
        L.ude_theta_size.restype = i64
        L.ude_theta_size.argtypes = [C.POINTER(UdeDesc)]
        L.ude_macs_per_eval.restype = i64
        L.ude_macs_per_eval.argtypes = [C.POINTER(UdeDesc)]
        L.ude_plan.restype = C.c_int
        L.ude_plan.argtypes = [C.POINTER(UdeDesc), C.c_int, C.c_int, C.POINTER(UdePlan)]
        L.ude_rk4_grid.restype = C.c_int
        L.ude_rk4_grid.argtypes = [vp, i32, f32, i32, i32, vp, C.POINTER(i32), vp, vp, vp]
        L.ude_rk4_forward.restype = C.c_int
        L.ude_rk4_forward.argtypes = [C.POINTER(UdeDesc), vp, vp, i32, vp, i32, vp, vp, vp, vp, vp, vp, vp, i64, vp]
        L.ude_rk4_backward.restype = C.c_int
        L.ude_rk4_backward.argtypes = [C.POINTER(UdeDesc), vp, vp, i32, vp, i32, vp, vp, vp, vp, vp, vp, vp, vp, vp,
                                       i64, vp]
        L.ude_rk4_forward_reparam.restype = C.c_int
        L.ude_rk4_forward_reparam.argtypes = [C.POINTER(UdeDesc), vp, C.POINTER(UdeReparam), i32, vp, i32, vp, vp, vp, vp, vp, vp,
                                              vp, i64, vp]
        L.ude_rk4_backward_reparam.restype = C.c_int
        L.ude_rk4_backward_reparam.argtypes = [C.POINTER(UdeDesc), vp, C.POINTER(UdeReparam), vp, i32, vp, i32, vp, vp, vp, vp, vp,
                                               vp, vp, vp, vp, vp, i64, vp]
        L.ude_rk4_forward_decoded.restype = C.c_int
        L.ude_rk4_forward_decoded.argtypes = [C.POINTER(UdeDesc), vp, vp, C.POINTER(UdeReparam), i32, vp, i32, vp, vp, vp,
                                              C.POINTER(UdeDecoder), vp, vp, vp, vp, vp, vp, i64, vp]
        L.ude_rk4_backward_decoded.restype = C.c_int
        L.ude_rk4_backward_decoded.argtypes = [C.POINTER(UdeDesc), vp, vp, C.POINTER(UdeReparam), i32, vp, i32, vp, vp, vp,
                                               C.POINTER(UdeDecoder), vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, i64, vp]
        L.ude_rk4_forward_bayes.restype = C.c_int
        L.ude_rk4_forward_bayes.argtypes = [C.POINTER(UdeDesc), vp, vp, i32, vp, i32, vp, vp, vp, vp, vp, vp, vp, i64, vp]
        L.ude_rk4_backward_bayes.restype = C.c_int
        L.ude_rk4_backward_bayes.argtypes = [C.POINTER(UdeDesc), vp, vp, vp, i32, vp, i32, vp, vp, vp, vp, vp, vp, vp, vp, vp, i64, vp]
        L.ude_rk4_forward_bayes_decoded.restype = C.c_int
        L.ude_rk4_forward_bayes_decoded.argtypes = [C.POINTER(UdeDesc), vp, vp, C.POINTER(UdeReparam), i32, vp, i32, vp, vp, vp,
                                                    C.POINTER(UdeDecoder), vp, vp, vp, vp, vp, vp, i64, vp]
        L.ude_rk4_backward_bayes_decoded.restype = C.c_int
        L.ude_rk4_backward_bayes_decoded.argtypes = [C.POINTER(UdeDesc), vp, vp, vp, C.POINTER(UdeReparam), i32, vp, i32, vp, vp, vp,
                                                     C.POINTER(UdeDecoder), vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, i64, vp]
        L.ude_dopri5_forward.restype = C.c_int
        L.ude_dopri5_forward.argtypes = [C.POINTER(UdeDesc), C.POINTER(UdeDopri5Opts), vp, vp, i32, vp, vp, vp, vp, vp,
                                         vp, i64, vp]
        L.ude_dopri5_backward.restype = C.c_int
        L.ude_dopri5_backward.argtypes = [C.POINTER(UdeDesc), C.POINTER(UdeDopri5Opts), vp, vp, i32, vp, vp, vp, vp, vp,
                                          vp, vp, vp, i64, vp]
        L.ude_rk4_path.restype = C.c_int
        L.ude_rk4_path.argtypes = [C.POINTER(UdeDesc)]
        L.ude_measure_ffma_tflops.restype = C.c_int
        L.ude_measure_ffma_tflops.argtypes = [C.POINTER(C.c_double), vp]
        if L.ude_abi_version() != 1:
            raise RuntimeError("libude_b200.so ABI version mismatch")
        _lib = L
    return _lib


class UdeError(RuntimeError):
    status = None       # the UdeStatus code when the error came from the C ABI


def check(status: int, what: str) -> None:
    if status != UDE_OK:
        L = lib()
        msg = L.ude_strerror(status).decode()
        if status == -4:
            msg += " — " + L.ude_last_cuda_error().decode()
        err = UdeError(f"{what}: {msg} (status {status})")
        err.status = status
        raise err

"""ctypes binding of libhdpde_b200.so (the C ABI declared in include/hdpde.h).

This is the Python twin of the Julia `ccall` glue in julia/HighDimPDEB200.jl:
it only marshals POD descriptors and raw pointers.  There is no CPU fallback:
if the shared library is missing, or no B200 is visible, every entry point
raises.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("HDPDE_B200_LIB") or os.path.join(_HERE, "libhdpde_b200.so")

# ---- status codes (include/hdpde.h) ----
HDP_OK = 0
HDP_ERR_BAD_ARG = -1
HDP_ERR_START_OUTSIDE_DOMAIN = -2
HDP_ERR_X0_SAMPLE_NOT_SUPPORTED = -3
HDP_ERR_CUDA = -4
HDP_ERR_NCCL = -5
HDP_ERR_NO_DEVICE = -6
HDP_ERR_ALLOC = -7
HDP_ERR_UNSUPPORTED = -8
HDP_ERR_NVRTC = -9

HDP_MAX_LEVELS = 6
HDP_PEER_HANDLE_BYTES, HDP_PEER_MAX_RANKS, HDP_PEER_SLOT_DOUBLES = 64, 16, 8192
HDP_MEM_HOST = 0
HDP_MEM_DEVICE = 1
HDP_ASYNC = 2

HDP_COEF_CONST, HDP_COEF_LINEAR, HDP_COEF_USER, HDP_COEF_QUADRATIC = 0, 1, 2, 3
HDP_G_USER, HDP_F_USER, HDP_F_NL_USER = 64, 64, 65
HDP_G_SUMSQ, HDP_G_INV_SUMSQ, HDP_G_EXP_SUMSQ, HDP_G_MIN, HDP_G_GAUSS = 0, 1, 2, 3, 4
HDP_F_ZERO, HDP_F_LINEAR, HDP_F_ALLEN_CAHN, HDP_F_DEFAULT_RISK, HDP_F_FISHER_KPP = 0, 1, 2, 3, 4
HDP_F_NL_FIRST = 16
HDP_F_NL_ALLEN_CAHN_DIFF, HDP_F_NL_FISHER, HDP_F_NL_REPLICATOR = 16, 17, 18
HDP_SAMPLER_NONE, HDP_SAMPLER_UNIFORM, HDP_SAMPLER_NORMAL = 0, 1, 2
HDP_NORMAL_DEFAULT, HDP_NORMAL_BM64, HDP_NORMAL_BM32 = 0, 1, 2
(HDP_STREAM_TIME, HDP_STREAM_STEP, HDP_STREAM_SAMPLER, HDP_STREAM_SDE, HDP_STREAM_X0,
 HDP_STREAM_MC) = range(6)

_DP = C.POINTER(C.c_double)


class hdp_sampler(C.Structure):
    _fields_ = [("kind", C.c_int32), ("shifted", C.c_int32), ("bounds_are_scalar", C.c_int32),
                ("_pad", C.c_int32), ("a_scalar", C.c_double), ("b_scalar", C.c_double),
                ("sigma", C.c_double), ("a", _DP), ("b", _DP)]


class hdp_problem(C.Structure):
    _fields_ = [("d", C.c_int32), ("mu_kind", C.c_int32), ("sigma_kind", C.c_int32),
                ("g_kind", C.c_int32), ("f_kind", C.c_int32), ("has_bc", C.c_int32),
                ("bc_is_scalar", C.c_int32), ("_pad", C.c_int32),
                ("t0", C.c_double), ("t1", C.c_double),
                ("mu_p", C.c_double * 2), ("sigma_p", C.c_double * 2),
                ("g_p", C.c_double * 4), ("f_p", C.c_double * 8),
                ("lo_scalar", C.c_double), ("hi_scalar", C.c_double),
                ("lo", _DP), ("hi", _DP)]


class hdp_user_exprs(C.Structure):
    _fields_ = [("mu", C.c_char_p), ("sigma", C.c_char_p), ("g", C.c_char_p), ("f", C.c_char_p)]


class hdp_mlp_cfg(C.Structure):
    _fields_ = [("M", C.c_int32), ("L", C.c_int32), ("K", C.c_int32), ("n_solves", C.c_int32),
                ("n_x0", C.c_int32), ("normal_mode", C.c_int32), ("compat_mlt_threads", C.c_int32),
                ("target_units", C.c_int32), ("seed", C.c_uint64), ("sampler", hdp_sampler),
                ("r_tape", C.c_void_p), ("n_tape", C.c_void_p), ("s_tape", C.c_void_p),
                ("user", C.POINTER(hdp_user_exprs)),
                ("group_size", C.c_int32), ("mem_flags", C.c_uint32), ("unit_oversub", C.c_double),
                ("std_err_out", C.c_void_p)]


class hdp_mlp_counts(C.Structure):
    _fields_ = [("path_steps", C.c_uint64), ("time_uniforms", C.c_uint64),
                ("sampler_calls", C.c_uint64), ("f_evals", C.c_uint64), ("g_evals", C.c_uint64)]


class hdp_mlp_stats(C.Structure):
    _fields_ = [("per_solve", hdp_mlp_counts), ("bounces", C.c_uint64), ("n_units", C.c_uint64),
                ("n_nodes", C.c_uint64), ("n_sites", C.c_uint64), ("device_ms", C.c_double),
                ("group_size", C.c_int32), ("n_launches", C.c_int32),
                ("phase_ms", C.c_double * 4),
                ("level_n", C.c_uint64 * (HDP_MAX_LEVELS + 1)), ("level_sum", C.c_double * (HDP_MAX_LEVELS + 1)),
                ("level_sumsq", C.c_double * (HDP_MAX_LEVELS + 1)), ("std_error", C.c_double),
                ("exchange_bytes", C.c_uint64), ("exchange_kind", C.c_int32), ("_pad", C.c_int32)]


class hdp_devinfo(C.Structure):
    _fields_ = [("device", C.c_int32), ("sm_count", C.c_int32), ("cc_major", C.c_int32),
                ("cc_minor", C.c_int32), ("clock_khz", C.c_int32), ("nccl_rank", C.c_int32),
                ("nccl_nranks", C.c_int32), ("_pad", C.c_int32), ("total_mem", C.c_uint64),
                ("name", C.c_char * 128)]


# every symbol include/hdpde.h declares: name -> (restype, argtypes)
_VP = C.c_void_p
_H = C.c_void_p
_PP = C.POINTER(hdp_problem)
_SP = C.POINTER(hdp_sampler)
_i, _u, _u64 = C.c_int, C.c_uint, C.c_uint64
SYMBOLS = {
    "hdp_create": (_i, [C.POINTER(_H), _i]),
    "hdp_destroy": (_i, [_H]),
    "hdp_last_error": (C.c_char_p, [_H]),
    "hdp_version": (_i, []),
    "hdp_device_info": (_i, [_H, C.POINTER(hdp_devinfo)]),
    "hdp_sync": (_i, [_H]),
    "hdp_nccl_unique_id": (_i, [C.POINTER(C.c_uint8)]),
    "hdp_comm_init_rank": (_i, [_H, C.POINTER(C.c_uint8), _i, _i]),
    "hdp_comm_destroy": (_i, [_H]),
    "hdp_comm_allreduce_f64": (_i, [_H, _VP, C.c_size_t]),
    "hdp_peer_window_export": (_i, [_H, C.POINTER(C.c_uint8)]),
    "hdp_peer_window_attach": (_i, [_H, C.POINTER(C.c_uint8), _i, _i]),
    "hdp_peer_window_detach": (_i, [_H]),
    "hdp_reflect_f64": (_i, [_H, _VP, _VP, _VP, _i, _i, _VP, _VP, _i, _VP, _u]),
    "hdp_reflect_f32": (_i, [_H, _VP, _VP, _VP, _i, _i, _VP, _VP, _i, _VP, _u]),
    "hdp_mc_sample_f64": (_i, [_H, _VP, _VP, _i, _i, _i, _SP, _u64, _VP, _u]),
    "hdp_mc_sample_f32": (_i, [_H, _VP, _VP, _i, _i, _i, _SP, _u64, _VP, _u]),
    "hdp_sde_loop_f64": (_i, [_H, _PP, _VP, _VP, _i, _i, C.c_double, _VP, _SP, _u64, _VP, _SP, _VP, _i, _i, _u]),
    "hdp_sde_loop_f32": (_i, [_H, _PP, _VP, _VP, _i, _i, C.c_float, _VP, _SP, _u64, _VP, _SP, _VP, _i, _i, _u]),
    "hdp_sde_loop_epochs_f64": (_i, [_H, _PP, _VP, _VP, _i, _i, C.c_double, _VP, _SP, _u64, _i, _SP, _VP, _i, _u]),
    "hdp_sde_loop_epochs_f32": (_i, [_H, _PP, _VP, _VP, _i, _i, C.c_float, _VP, _SP, _u64, _i, _SP, _VP, _i, _u]),
    "hdp_mlt_sde_loop_f64": (_i, [_H, _PP, _VP, _VP, _i, C.c_double, C.c_double, _u64, _VP, _u]),
    "hdp_mlt_sde_loop_f32": (_i, [_H, _PP, _VP, _VP, _i, C.c_float, C.c_float, _u64, _VP, _u]),
    "hdp_set_user_exprs": (_i, [_H, C.POINTER(hdp_user_exprs)]),
    "hdp_em_ensemble_f64": (_i, [_H, _PP, _VP, _i, _i, C.c_double, C.c_double, _VP, _u64, _VP, _u]),
    "hdp_em_ensemble_f32": (_i, [_H, _PP, _VP, _i, _i, C.c_float, C.c_float, _VP, _u64, _VP, _u]),
    "hdp_mlp_counts_get": (_i, [_i, _i, _i, _i, _i, C.POINTER(hdp_mlp_counts)]),
    "hdp_mlp_solve_f64": (_i, [_H, _PP, C.POINTER(hdp_mlp_cfg), _VP, _VP, C.POINTER(hdp_mlp_stats)]),
    "hdp_mlp_solve_f32": (_i, [_H, _PP, C.POINTER(hdp_mlp_cfg), _VP, _VP, C.POINTER(hdp_mlp_stats)]),
    "hdp_user_kinds_check": (_i, [_PP, C.POINTER(hdp_user_exprs), _i, C.c_char_p, C.c_size_t]),
    "hdp_user_coefs_check": (_i, [_PP, C.POINTER(hdp_user_exprs), _i, C.c_char_p, C.c_size_t]),
    "hdp_tape_fill_f64": (_i, [_H, _VP, _i, _i, _i, _u64, _u64, _u64, _i]),
    "hdp_tape_fill_f32": (_i, [_H, _VP, _i, _i, _i, _u64, _u64, _u64, _i]),
    "hdp_measure_pipe_peak": (_i, [_H, _i, C.POINTER(C.c_double)]),
    "hdp_last_device_ms": (_i, [_H, C.POINTER(C.c_double)]),
}


class HdpError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"libhdpde_b200 error {code}: {msg}")
        self.code = code


class StartOutsideDomain(HdpError):
    """`_reflect`: 'a not in hypercube' (src/reflect.jl:12, :59)."""


_lib = None
_lib_lock = threading.Lock()


def load_library(path: str | None = None):
    """dlopen libhdpde_b200.so and bind every declared symbol.  Fails loudly."""
    global _lib
    with _lib_lock:
        if _lib is not None and path is None:
            return _lib
        p = path or LIB_PATH
        if not os.path.exists(p):
            raise ImportError(
                f"{p} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                f"or `make -C {os.path.join(_HERE, 'csrc')}`.  There is no CPU fallback.")
        lib = C.CDLL(p)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(lib, name)  # AttributeError if a declared symbol is missing
            fn.restype = res
            fn.argtypes = args
        _lib = lib
        return lib


def check(handle, rc: int):
    if rc == HDP_OK:
        return
    lib = load_library()
    msg = lib.hdp_last_error(handle)
    msg = msg.decode("utf-8", "replace") if msg else ""
    if rc == HDP_ERR_START_OUTSIDE_DOMAIN:
        raise StartOutsideDomain(rc, msg)
    raise HdpError(rc, msg)


def np_ptr(a: np.ndarray | None):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Handle:
    """One handle per host task; owns a CUDA stream, scratch and (optionally) an
    NCCL communicator (include/hdpde.h: hdp_create)."""

    def __init__(self, device: int = -1):
        self.lib = load_library()
        h = C.c_void_p()
        rc = self.lib.hdp_create(C.byref(h), int(device))
        if rc != HDP_OK:
            msg = self.lib.hdp_last_error(None)
            raise HdpError(rc, msg.decode() if msg else "hdp_create failed")
        self.h = h
        self._keep = []

    def close(self):
        if getattr(self, "h", None):
            self.lib.hdp_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def info(self) -> dict:
        di = hdp_devinfo()
        check(self.h, self.lib.hdp_device_info(self.h, C.byref(di)))
        return {"device": di.device, "sm_count": di.sm_count, "cc": (di.cc_major, di.cc_minor),
                "clock_khz": di.clock_khz, "nccl_rank": di.nccl_rank, "nccl_nranks": di.nccl_nranks,
                "total_mem": di.total_mem, "name": di.name.decode()}

    def sync(self):
        check(self.h, self.lib.hdp_sync(self.h))

    # ---- NCCL plumbing ----
    @staticmethod
    def nccl_unique_id() -> bytes:
        lib = load_library()
        buf = (C.c_uint8 * 128)()
        rc = lib.hdp_nccl_unique_id(buf)
        if rc != HDP_OK:
            msg = lib.hdp_last_error(None)
            raise HdpError(rc, msg.decode() if msg else "")
        return bytes(buf)

    def comm_init_rank(self, uid: bytes, nranks: int, rank: int):
        buf = (C.c_uint8 * 128).from_buffer_copy(uid)
        check(self.h, self.lib.hdp_comm_init_rank(self.h, buf, nranks, rank))

    def comm_destroy(self):
        check(self.h, self.lib.hdp_comm_destroy(self.h))

    # ---- peer exchange windows (hdp_peer_window_*) ----
    def peer_window_export(self) -> bytes:
        buf = (C.c_uint8 * HDP_PEER_HANDLE_BYTES)()
        check(self.h, self.lib.hdp_peer_window_export(self.h, buf))
        return bytes(buf)

    def peer_window_attach(self, handles: bytes, nranks: int, rank: int):
        assert len(handles) == nranks * HDP_PEER_HANDLE_BYTES
        buf = (C.c_uint8 * len(handles)).from_buffer_copy(handles)
        check(self.h, self.lib.hdp_peer_window_attach(self.h, buf, nranks, rank))

    def peer_window_detach(self):
        check(self.h, self.lib.hdp_peer_window_detach(self.h))

    def allreduce_f64(self, a: np.ndarray) -> np.ndarray:
        a = np.ascontiguousarray(a, dtype=np.float64)
        check(self.h, self.lib.hdp_comm_allreduce_f64(self.h, np_ptr(a), a.size))
        return a

    def measure_pipe_peak(self, which: int) -> float:
        out = C.c_double()
        check(self.h, self.lib.hdp_measure_pipe_peak(self.h, which, C.byref(out)))
        return out.value

    def last_device_ms(self) -> float:
        out = C.c_double()
        check(self.h, self.lib.hdp_last_device_ms(self.h, C.byref(out)))
        return out.value

    def tape_fill(self, dtype, stream: int, want_normal: bool, seed: int, ordinal0: int, n_ord: int,
                  width: int, normal_mode: int = HDP_NORMAL_DEFAULT) -> np.ndarray:
        dtype = np.dtype(dtype)
        out = np.empty((n_ord, width), dtype=dtype)
        fn = self.lib.hdp_tape_fill_f64 if dtype == np.float64 else self.lib.hdp_tape_fill_f32
        check(self.h, fn(self.h, np_ptr(out), stream, int(want_normal), normal_mode, seed, ordinal0, n_ord, width))
        return out


_default = None
_default_lock = threading.Lock()


def default_handle() -> Handle:
    global _default
    with _default_lock:
        if _default is None:
            _default = Handle(-1)
        return _default


def set_default_handle(h: Handle | None):
    global _default
    with _default_lock:
        _default = h

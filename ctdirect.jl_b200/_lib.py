"""ctypes binding of libctdirect_b200.so (include/ctdirect_b200.h).  No fallback: if the CUDA library is
missing this raises, and evaluation calls raise when no GPU is present."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("CTD_LIB") or os.path.join(HERE, "lib", "libctdirect_b200.so")

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int64)


class Desc(C.Structure):
    _fields_ = [
        ("problem", C.c_char_p), ("scheme", C.c_char_p), ("grid_size", C.c_int32), ("time_grid", _dp),
        ("n_time_grid", C.c_int32), ("params", _dp), ("n_params", C.c_int32), ("pattern_mode", C.c_int32),
        ("batch", C.c_int32), ("device", C.c_int32), ("control_steps", C.c_int32),
    ]


class Solution(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("time", "state", "control", "variable", "costate", "path_dual",
                                          "boundary_dual")]


class CtdError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"[ctdirect_b200 {code}] {msg}")
        self.code = code
        self.msg = msg


_lib = None

# name -> (restype, argtypes); every symbol declared in include/ctdirect_b200.h
_H = C.c_void_p
SIGNATURES = {
    "ctd_create": (C.c_int, [C.POINTER(Desc), C.POINTER(_H)]),
    "ctd_destroy": (C.c_int, [_H]),
    "ctd_last_error": (C.c_char_p, []),
    "ctd_list_models": (C.c_int64, [C.c_char_p, C.c_int64]),
    "ctd_dims": (C.c_int, [_H, _ip, _ip, _ip, _ip]),
    "ctd_info": (C.c_int, [_H, _ip]),
    "ctd_bounds": (C.c_int, [_H, _dp, _dp, _dp, _dp]),
    "ctd_x0": (C.c_int, [_H, _dp]),
    "ctd_time_grid": (C.c_int, [_H, _dp]),
    "ctd_jac_structure": (C.c_int, [_H, _ip, _ip]),
    "ctd_hess_structure": (C.c_int, [_H, _ip, _ip]),
    "ctd_obj": (C.c_int, [_H, C.c_void_p, C.c_void_p]),
    "ctd_grad": (C.c_int, [_H, C.c_void_p, C.c_void_p]),
    "ctd_cons": (C.c_int, [_H, C.c_void_p, C.c_void_p]),
    "ctd_jac_coord": (C.c_int, [_H, C.c_void_p, C.c_void_p]),
    "ctd_hess_coord": (C.c_int, [_H, C.c_void_p, C.c_void_p, C.c_double, C.c_void_p]),
    "ctd_eval": (C.c_int, [_H, C.c_int, C.c_void_p, C.c_void_p, C.c_double] + [C.c_void_p] * 5),
    "ctd_eval_async": (C.c_int, [_H, C.c_int, C.c_void_p, C.c_void_p, C.c_double] + [C.c_void_p] * 5),
    "ctd_eval_dev": (C.c_int, [_H, C.c_int, C.c_void_p, C.c_void_p, C.c_double] + [C.c_void_p] * 5),
    "ctd_peer_init": (C.c_int, [_H]),
    "ctd_eval_gather_dev": (C.c_int, [_H, C.c_int, C.c_void_p, C.c_void_p, C.c_double] + [C.c_void_p] * 4 + [C.c_int32]),
    "ctd_peer_wait": (C.c_int, [_H, C.c_int32, C.c_int32, C.POINTER(C.c_void_p)]),
    "ctd_eval_norms_dev": (C.c_int, [_H, C.c_int, C.c_void_p, C.c_void_p, C.c_double] + [C.c_void_p] * 6),
    "ctd_norms_dev": (C.c_int, [_H, C.c_void_p, C.c_void_p, C.c_void_p]),
    "ctd_violation_dev": (C.c_int, [_H, C.c_void_p, C.c_void_p]),
    "ctd_norms": (C.c_int, [_H, C.c_void_p, C.c_void_p, C.c_void_p]),
    "ctd_unpack": (C.c_int, [_H, C.c_void_p, C.c_void_p, C.POINTER(Solution)]),
    "ctd_unpack_dev": (C.c_int, [_H, C.c_void_p, C.c_void_p, C.POINTER(Solution)]),
    "ctd_control_steps": (C.c_int32, [_H]),
    "ctd_control_time": (C.c_int, [_H, _dp, _dp]),
    "ctd_pack": (C.c_int, [_H] + [C.c_void_p] * 5),
    "ctd_pack_dev": (C.c_int, [_H] + [C.c_void_p] * 5),
    "ctd_set_params": (C.c_int, [_H, _dp, C.c_int32]),
    "ctd_sync": (C.c_int, [_H]),
    "ctd_stream": (C.c_void_p, [_H]),
    "ctd_set_stream": (C.c_int, [_H, C.c_void_p]),
    "ctd_launch_count": (C.c_int64, [_H]),
    "ctd_counters": (C.c_int, [_H, C.POINTER(C.c_int64)]),
    "ctd_reset_counters": (C.c_int, [_H]),
    "ctd_stats_json": (C.c_int64, [_H, C.c_char_p, C.c_int64]),
    "ctd_kkt_nnz": (C.c_int64, [_H]),
    "ctd_kkt_structure": (C.c_int, [_H, _ip, _ip]),
    "ctd_kkt_assemble_dev": (C.c_int, [_H, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "ctd_nccl_unique_id": (C.c_int, [C.c_char_p]),
    "ctd_nccl_version": (C.c_int, []),
    "ctd_comm_init": (C.c_int, [_H, C.c_char_p, C.c_int32, C.c_int32]),
    "ctd_gather_norms_dev": (C.c_int, [_H, C.c_void_p, C.c_void_p, C.c_int32]),
    "ctd_comm_wait": (C.c_int, [_H, C.c_int32]),
}


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise CtdError(-100, f"{LIB_PATH} not found: build it with `python -m ctdirect_b200.build` "
                             "(nvcc, sm_100a).  There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(code):
    if code != 0:
        raise CtdError(code, load().ctd_last_error().decode())


def list_models():
    lib = load()
    n = lib.ctd_list_models(None, 0)
    buf = C.create_string_buffer(int(n))
    lib.ctd_list_models(buf, n)
    return [tuple(l.split()) for l in buf.value.decode().splitlines()]

"""TEST INFRASTRUCTURE ONLY -- ctypes wrapper around the CPU oracle (oracle/_build/liboracle.so).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  The product package never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "liboracle.so")
_lib = None

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int64)


def build(force: bool = False) -> str:
    """Compile the oracle (gcc only; no GPU, no reference sources needed)."""
    srcs = [os.path.join(_HERE, f) for f in ("oracle_capi.cpp", "docp_oracle.hpp", "problems.hpp", "hdual.hpp", "tracer.hpp")]
    if (
        force
        or not os.path.exists(_LIB_PATH)
        or os.path.getmtime(_LIB_PATH) < max(os.path.getmtime(s) for s in srcs)
    ):
        subprocess.check_call(["make", "-C", _HERE], stdout=subprocess.DEVNULL)
    return _LIB_PATH


def _load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIB_PATH):
        build()
    lib = C.CDLL(_LIB_PATH)
    lib.orc_create.restype = C.c_void_p
    lib.orc_create.argtypes = [C.c_char_p, C.c_char_p, C.c_int, _dp, C.c_int, _dp, C.c_int, C.c_int]
    lib.orc_last_error.restype = C.c_char_p
    lib.orc_destroy.argtypes = [C.c_void_p]
    lib.orc_dims.argtypes = [C.c_void_p, _ip]
    lib.orc_flags.argtypes = [C.c_void_p, C.POINTER(C.c_int)]
    lib.orc_jac_structure.argtypes = [C.c_void_p, _ip, _ip]
    lib.orc_hess_structure.argtypes = [C.c_void_p, _ip, _ip]
    lib.orc_bounds.argtypes = [C.c_void_p, _dp, _dp, _dp, _dp]
    lib.orc_x0.argtypes = [C.c_void_p, _dp]
    lib.orc_time_grid.argtypes = [C.c_void_p, _dp, _dp]
    lib.orc_obj.restype = C.c_double
    lib.orc_obj.argtypes = [C.c_void_p, _dp]
    lib.orc_cons.argtypes = [C.c_void_p, _dp, _dp]
    lib.orc_derivs.argtypes = [C.c_void_p, _dp, _dp, C.c_double, _dp, _dp, _dp, _ip]
    lib.orc_unpack.argtypes = [C.c_void_p] + [_dp] * 9
    lib.orc_pack.argtypes = [C.c_void_p] + [_dp] * 5
    lib.orc_control_steps.argtypes = [C.c_void_p]
    lib.orc_control_time.argtypes = [C.c_void_p, _dp, _dp]
    lib.orc_set_threads.restype = C.c_int
    lib.orc_set_threads.argtypes = [C.c_int]
    _lib = lib
    return lib


def _p(a):
    return None if a is None else a.ctypes.data_as(_dp)


def set_threads(n: int) -> int:
    return _load().orc_set_threads(int(n))


class OracleModel:
    """CPU oracle for one discretised OCP (mirrors the NLPModels callbacks the reference exposes)."""

    def __init__(self, problem: str, scheme: str = "midpoint", grid_size: int = 250, time_grid=None,
                 params=None, pattern_mode: str = "reference"):
        lib = _load()
        tg = None if time_grid is None else np.ascontiguousarray(time_grid, dtype=np.float64)
        par = None if params is None else np.ascontiguousarray(params, dtype=np.float64)
        pm = {"reference": 0, "corrected": 1, "optimized": 2}[pattern_mode]
        self._h = lib.orc_create(problem.encode(), scheme.encode(), int(grid_size), _p(tg),
                                 0 if tg is None else len(tg), _p(par), 0 if par is None else len(par), pm)
        if not self._h:
            raise ValueError(lib.orc_last_error().decode())
        self._lib = lib
        d = np.zeros(12, dtype=np.int64)
        lib.orc_dims(self._h, d.ctypes.data_as(_ip))
        (self.nvar, self.ncon, self.nnzj, self.nnzh, self.nx, self.nu, self.nv, self.np_, self.nb,
         self.N, self.step_block, self.cblock) = (int(v) for v in d)
        f = (C.c_int * 5)()
        lib.orc_flags(self._h, f)
        self.free_t0, self.free_tf, self.has_lagrange, self.has_mayer, self.is_max = (bool(v) for v in f)
        self.minimize = not self.is_max
        self.control_steps = int(lib.orc_control_steps(self._h))   # > 1: the DirectShooting discretizer (midpoint_cs*)

    def __del__(self):
        if getattr(self, "_h", None):
            self._lib.orc_destroy(self._h)
            self._h = None

    def jac_structure(self):
        r = np.zeros(self.nnzj, dtype=np.int64)
        c = np.zeros(self.nnzj, dtype=np.int64)
        self._lib.orc_jac_structure(self._h, r.ctypes.data_as(_ip), c.ctypes.data_as(_ip))
        return r, c

    def hess_structure(self):
        r = np.zeros(self.nnzh, dtype=np.int64)
        c = np.zeros(self.nnzh, dtype=np.int64)
        self._lib.orc_hess_structure(self._h, r.ctypes.data_as(_ip), c.ctypes.data_as(_ip))
        return r, c

    def bounds(self):
        lvar, uvar = np.zeros(self.nvar), np.zeros(self.nvar)
        lcon, ucon = np.zeros(self.ncon), np.zeros(self.ncon)
        self._lib.orc_bounds(self._h, _p(lvar), _p(uvar), _p(lcon), _p(ucon))
        return lvar, uvar, lcon, ucon

    def x0(self):
        x = np.zeros(self.nvar)
        self._lib.orc_x0(self._h, _p(x))
        return x

    def time_grid(self, x):
        x = np.ascontiguousarray(x, dtype=np.float64)
        t = np.zeros(self.N + 1)
        self._lib.orc_time_grid(self._h, _p(x), _p(t))
        return t

    def obj(self, x):
        x = np.ascontiguousarray(x, dtype=np.float64)
        return float(self._lib.orc_obj(self._h, _p(x)))

    def cons(self, x):
        x = np.ascontiguousarray(x, dtype=np.float64)
        c = np.zeros(self.ncon)
        self._lib.orc_cons(self._h, _p(x), _p(c))
        return c

    def derivs(self, x, y=None, obj_weight=1.0, want=("grad", "jac", "hess")):
        x = np.ascontiguousarray(x, dtype=np.float64)
        y = None if y is None else np.ascontiguousarray(y, dtype=np.float64)
        g = np.zeros(self.nvar) if "grad" in want else None
        j = np.zeros(self.nnzj) if "jac" in want else None
        h = np.zeros(self.nnzh) if "hess" in want else None
        dropped = np.zeros(2, dtype=np.int64)
        self._lib.orc_derivs(self._h, _p(x), _p(y), float(obj_weight), _p(g), _p(j), _p(h),
                             dropped.ctypes.data_as(_ip))
        self.dropped = (int(dropped[0]), int(dropped[1]))
        return g, j, h

    def unpack(self, x, y=None):
        """Solution arrays of build_OCP_solution (src/DOCP_data.jl:514-633), [node][component]."""
        x = np.ascontiguousarray(x, dtype=np.float64)
        y = None if y is None else np.ascontiguousarray(y, dtype=np.float64)
        N = self.N
        out = {"time": np.zeros(N + 1), "state": np.zeros((N + 1, self.nx)),
               "control": np.zeros((N * self.control_steps + 1, self.nu)),
               "variable": np.zeros(self.nv), "costate": np.zeros((N, self.nx)),
               "path_dual": np.zeros((N + 1, self.np_)), "boundary_dual": np.zeros(self.nb)}
        self._lib.orc_unpack(self._h, _p(x), _p(y), *[_p(out[k]) for k in
                             ("time", "state", "control", "variable", "costate", "path_dual", "boundary_dual")])
        out["control_time"] = np.zeros(N * self.control_steps + 1)
        self._lib.orc_control_time(self._h, _p(x), _p(out["control_time"]))
        return out

    def pack(self, state=None, control=None, variable=None, stage_control=None):
        """NLP vector of __initial_guess (DOCP_variables.jl:122-145) from sampled state / control / variable."""
        c = lambda a: None if a is None else np.ascontiguousarray(a, dtype=np.float64)
        state, control, variable, stage_control = c(state), c(control), c(variable), c(stage_control)
        x = np.zeros(self.nvar)
        self._lib.orc_pack(self._h, _p(state), _p(control), _p(variable), _p(stage_control), _p(x))
        return x

    def grad(self, x):
        return self.derivs(x, want=("grad",))[0]

    def jac_coord(self, x):
        return self.derivs(x, want=("jac",))[1]

    def hess_coord(self, x, y, obj_weight=1.0):
        return self.derivs(x, y, obj_weight, want=("hess",))[2]

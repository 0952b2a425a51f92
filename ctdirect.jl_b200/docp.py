"""Host-side mirror of the reference's interface for the DOCP evaluation path.

Names follow CTDirect.jl (paths relative to the reference tree):
  Collocation(grid_size, scheme, time_grid)     src/collocation.jl:8-48   (defaults 250 / :midpoint / nothing)
  discretize(ocp, discretizer) -> DOCP          src/CTDirect.jl:31-50, src/collocation.jl:57-73
  DOCP.nlp_model() / adnlp_model_builder(...)   src/collocation.jl:90-153 (returns the NLPModels-style model)
  direct_transcription / nlp_model / DOCP_*     legacy spellings (test/tmp/percival.jl:24, test/benchmark.jl:81-125)
and, for the model, the NLPModels API [3P] that Ipopt / MadNLP drive:
  obj, grad, cons, jac_structure, jac_coord, hess_structure, hess_coord(x, y; obj_weight), meta.

Arrays may be numpy arrays (host; copied through the handle, synchronous) or torch CUDA tensors
(device; no copies).  Device work is enqueued on torch's CURRENT stream by default (`stream="torch"`), so
it is ordered with the torch kernels that produce the inputs and consume the results; `stream="own"`
keeps the handle's private non-blocking stream (the caller then synchronises explicitly: `synchronize()`,
events, or `torch.cuda.ExternalStream(docp.cuda_stream)`).  All compute happens in libctdirect_b200.so.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Optional

import numpy as np

from . import _lib

CONS, JAC, HESS, GRAD, OBJ = 1, 2, 4, 8, 16

SCHEMES = (
    "trapeze", "midpoint", "euler", "euler_implicit", "gauss_legendre_2", "gauss_legendre_3",
    "gauss_legendre_1_constant_control", "gauss_legendre_2_constant_control", "gauss_legendre_3_constant_control",
)


@dataclass
class Collocation:
    """Collocation discretizer options (src/collocation.jl:20-42; alias disc_method for scheme)."""
    grid_size: int = 250
    scheme: str = "midpoint"
    time_grid: Optional[np.ndarray] = None

    def __post_init__(self):
        self.scheme = str(self.scheme).lstrip(":")
        if self.time_grid is not None:
            self.time_grid = np.ascontiguousarray(self.time_grid, dtype=np.float64)


@dataclass
class DirectShooting:
    """DirectShooting discretizer options (src/direct_shooting.jl:12-41): `control_steps` controls per time step,
    midpoint rule.  `discretize(ocp, DirectShooting(grid_size=50, control_steps=10))`."""
    grid_size: int = 250
    control_steps: int = 1
    scheme: str = "midpoint"
    time_grid = None            # DOCP(ocp, grid_size, control_steps, scheme, nothing): direct_shooting.jl:63-64

    def __post_init__(self):
        self.scheme = str(self.scheme).lstrip(":")


@dataclass
class Meta:
    """Subset of NLPModels' NLPModelMeta that CTDirect's builders fill (src/collocation.jl:137-149)."""
    nvar: int
    ncon: int
    nnzj: int
    nnzh: int
    x0: np.ndarray
    lvar: np.ndarray
    uvar: np.ndarray
    lcon: np.ndarray
    ucon: np.ndarray
    minimize: bool
    name: str


def _is_torch(a):
    return type(a).__module__.startswith("torch")


def _ptr(a):
    if a is None:
        return None
    if _is_torch(a):
        return C.c_void_p(a.data_ptr())
    return C.c_void_p(a.ctypes.data)


class DOCP:
    """Discretised optimal control problem bound to one GPU (DOCP struct, src/DOCP_data.jl:265-366)."""

    def __init__(self, ocp, grid_size: int = 250, scheme: str = "midpoint", time_grid=None, params=None,
                 pattern_mode: str = "reference", batch: int = 1, device: int = 0, disc_method: Optional[str] = None,
                 stream: str = "torch", control_steps: int = 1):
        if stream not in ("torch", "own"):
            raise ValueError("stream must be 'torch' (follow torch's current stream) or 'own'")
        self._stream_mode, self._stream_set = stream, None
        lib = _lib.load()
        name = ocp if isinstance(ocp, str) else ocp.name
        scheme = str(disc_method if disc_method is not None else scheme).lstrip(":")
        tg = None if time_grid is None else np.ascontiguousarray(time_grid, dtype=np.float64)
        par = None if params is None else np.ascontiguousarray(params, dtype=np.float64).ravel()
        d = _lib.Desc()
        self._keep = (name.encode(), scheme.encode(), tg, par)
        d.problem, d.scheme = self._keep[0], self._keep[1]
        d.grid_size = int(grid_size)
        d.time_grid = None if tg is None else tg.ctypes.data_as(_lib._dp)
        d.n_time_grid = 0 if tg is None else len(tg)
        d.params = None if par is None else par.ctypes.data_as(_lib._dp)
        d.n_params = 0 if par is None else len(par)
        d.pattern_mode = {"reference": 0, "corrected": 1, "optimized": 2}[pattern_mode]
        d.batch = int(batch)
        d.device = int(device)
        d.control_steps = int(control_steps)
        if int(control_steps) > 1:
            pattern_mode = "optimized"     # the only pattern DirectShooting has (direct_shooting.jl:101-137)
        h = C.c_void_p()
        _lib.check(lib.ctd_create(C.byref(d), C.byref(h)))
        self._h = h
        self._lib = lib
        self.problem, self.scheme, self.pattern_mode, self.device = name, scheme, pattern_mode, int(device)
        v = [C.c_int64() for _ in range(4)]
        _lib.check(lib.ctd_dims(h, *[C.byref(q) for q in v]))
        self.dim_NLP_variables, self.dim_NLP_constraints, self.nnzj, self.nnzh = (int(q.value) for q in v)
        info = np.zeros(16, dtype=np.int64)
        _lib.check(lib.ctd_info(h, info.ctypes.data_as(_lib._ip)))
        (self.NLP_x, self.NLP_u, self.NLP_v, self.path_cons, self.boundary_cons, self.steps, self.step_variables_block,
         self.step_constraints_block, self.batch) = (int(q) for q in info[:9])
        self.free_t0, self.free_tf, self.lagrange, self.mayer, self.max = (bool(q) for q in info[9:14])
        self.npar, self.stage = int(info[14]), int(info[15])
        self.control_steps = int(lib.ctd_control_steps(h))
        self._nlp = None

    # convenient aliases
    nvar = property(lambda self: self.dim_NLP_variables)
    ncon = property(lambda self: self.dim_NLP_constraints)

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            self._lib.ctd_destroy(h)
            self._h = None

    # ---- host-only structural queries ------------------------------------------------------------
    def bounds(self):
        lvar, uvar = np.empty(self.nvar), np.empty(self.nvar)
        lcon, ucon = np.empty(self.ncon), np.empty(self.ncon)
        _lib.check(self._lib.ctd_bounds(self._h, *[a.ctypes.data_as(_lib._dp) for a in (lvar, uvar, lcon, ucon)]))
        return lvar, uvar, lcon, ucon

    def initial_guess(self):
        x0 = np.empty(self.nvar)
        _lib.check(self._lib.ctd_x0(self._h, x0.ctypes.data_as(_lib._dp)))
        return x0

    def normalized_time_grid(self):
        s = np.empty(self.steps + 1)
        _lib.check(self._lib.ctd_time_grid(self._h, s.ctypes.data_as(_lib._dp)))
        return s

    def jacobian_structure(self):
        r, c = np.empty(self.nnzj, dtype=np.int64), np.empty(self.nnzj, dtype=np.int64)
        _lib.check(self._lib.ctd_jac_structure(self._h, r.ctypes.data_as(_lib._ip), c.ctypes.data_as(_lib._ip)))
        return r, c

    def hessian_structure(self):
        r, c = np.empty(self.nnzh, dtype=np.int64), np.empty(self.nnzh, dtype=np.int64)
        _lib.check(self._lib.ctd_hess_structure(self._h, r.ctypes.data_as(_lib._ip), c.ctypes.data_as(_lib._ip)))
        return r, c

    # ---- KKT assembly (ctd_kkt_* of the C ABI) ------------------------------------------------------
    @property
    def nnzk(self) -> int:
        return int(self._lib.ctd_kkt_nnz(self._h))

    def kkt_structure(self):
        """(colptr, rows), 1-based, of the lower triangle of [[H + diag(dx), .], [J, -diag(dc)]] in CSC order."""
        colptr = np.empty(self.nvar + self.ncon + 1, dtype=np.int64)
        rows = np.empty(self.nnzk, dtype=np.int64)
        _lib.check(self._lib.ctd_kkt_structure(self._h, colptr.ctypes.data_as(_lib._ip), rows.ctypes.data_as(_lib._ip)))
        return colptr, rows

    def kkt_assemble(self, hvals, jvals, dx=None, dc=None, out=None):
        """Move the value arrays of an evaluation (CUDA tensors) to KKT order in one launch."""
        hvals = self._check_in(hvals, self.nnzh, "hvals")
        jvals = self._check_in(jvals, self.nnzj, "jvals")
        dx = self._check_in(dx, self.nvar, "dx")
        dc = self._check_in(dc, self.ncon, "dc")
        if not _is_torch(hvals):
            raise TypeError("kkt_assemble needs CUDA tensors")
        out = self._check_in(out, self.nnzk, "out", output=True) if out is not None else self._empty_like(hvals, self.nnzk)
        self._follow_stream(hvals)
        _lib.check(self._lib.ctd_kkt_assemble_dev(self._h, _ptr(hvals), _ptr(jvals), _ptr(dx), _ptr(dc), _ptr(out)))
        return out

    # ---- evaluation --------------------------------------------------------------------------------
    def _shape(self, n):
        if n == 1:
            return (self.batch,)          # one objective value per instance
        return (n,) if self.batch == 1 else (self.batch, n)

    def _empty_like(self, x, n):
        if _is_torch(x):
            import torch
            return torch.empty(self._shape(n), dtype=torch.float64, device=x.device)
        return np.empty(self._shape(n))

    def _check_in(self, a, n, what, output=False):
        """Inputs: CUDA tensors must be contiguous float64 on the handle's device, host arrays are converted.
        Outputs (`output=True`) are filled IN PLACE (NLPModels semantics), so a host buffer that would have to
        be copied to become C-contiguous float64 is an error, not a silent temporary."""
        if a is None:
            return None
        if _is_torch(a):
            import torch
            if a.dtype != torch.float64 or not a.is_cuda or not a.is_contiguous():
                raise TypeError(f"{what}: expected a contiguous float64 CUDA tensor")
            if a.device.index != self.device:
                raise ValueError(f"{what}: tensor is on cuda:{a.device.index}, handle on cuda:{self.device}")
        elif output:
            if not (isinstance(a, np.ndarray) and a.dtype == np.float64 and a.flags.c_contiguous and a.flags.writeable):
                raise TypeError(f"{what}: an output buffer must be a writeable C-contiguous float64 numpy array "
                                "(it is filled in place)")
        else:
            a = np.ascontiguousarray(a, dtype=np.float64)
        size = a.numel() if _is_torch(a) else a.size
        if size != self.batch * n:
            raise ValueError(f"{what}: expected {self.batch}x{n} values, got {size}")
        return a

    def _follow_stream(self, *tensors):
        """`stream="torch"`: put the handle on torch's current stream of its device (a no-op when it already is
        there), so that device evaluations are ordered with the surrounding torch work."""
        if self._stream_mode != "torch" or not any(_is_torch(t) for t in tensors if t is not None):
            return
        import torch
        ptr = int(torch.cuda.current_stream(self.device).cuda_stream) or 1   # 0 = legacy default stream = cudaStreamLegacy (1)
        if ptr != self._stream_set:
            _lib.check(self._lib.ctd_set_stream(self._h, C.c_void_p(ptr)))
            self._stream_set = ptr

    def evaluate(self, what: int, x, y=None, obj_weight: float = 1.0, out=None):
        """One fused kernel launch for the outputs selected by `what` (CONS|JAC|HESS|GRAD|OBJ subset).
        Returns a dict with keys among 'obj', 'grad', 'cons', 'jac', 'hess'."""
        x = self._check_in(x, self.nvar, "x")
        y = self._check_in(y, self.ncon, "y")
        dev = _is_torch(x)
        if dev and y is not None and not _is_torch(y):
            raise TypeError("x and y must both be host arrays or both CUDA tensors")
        out = dict(out or {})
        res = {}
        for bit, key, n in ((OBJ, "obj", 1), (GRAD, "grad", self.nvar), (CONS, "cons", self.ncon),
                            (JAC, "jac", self.nnzj), (HESS, "hess", self.nnzh)):
            if what & bit:
                buf = out.get(key)
                res[key] = self._check_in(buf, n, key, output=True) if buf is not None else self._empty_like(x, n)
        self._follow_stream(x)
        fn = self._lib.ctd_eval_dev if dev else self._lib.ctd_eval
        _lib.check(fn(self._h, int(what), _ptr(x), _ptr(y), float(obj_weight), _ptr(res.get("obj")),
                      _ptr(res.get("grad")), _ptr(res.get("cons")), _ptr(res.get("jac")), _ptr(res.get("hess"))))
        return res

    def evaluate_with_norms(self, what: int, x, y=None, obj_weight: float = 1.0, out=None, viol=None):
        """`evaluate` on CUDA tensors plus the per-instance constraint violation, reduced in the same launch.
        Returns (result dict, viol tensor)."""
        import torch
        x = self._check_in(x, self.nvar, "x")
        y = self._check_in(y, self.ncon, "y")
        if not _is_torch(x) or not (what & CONS):
            raise TypeError("evaluate_with_norms needs CUDA tensors and CONS among the outputs")
        out = dict(out or {})
        res = {}
        for bit, key, n in ((OBJ, "obj", 1), (GRAD, "grad", self.nvar), (CONS, "cons", self.ncon),
                            (JAC, "jac", self.nnzj), (HESS, "hess", self.nnzh)):
            if what & bit:
                buf = out.get(key)
                res[key] = self._check_in(buf, n, key, output=True) if buf is not None else self._empty_like(x, n)
        if viol is None:
            viol = torch.empty(self.batch, dtype=torch.float64, device=x.device)
        self._follow_stream(x)
        _lib.check(self._lib.ctd_eval_norms_dev(self._h, int(what), _ptr(x), _ptr(y), float(obj_weight),
                                                _ptr(res.get("obj")), _ptr(res.get("grad")), _ptr(res.get("cons")),
                                                _ptr(res.get("jac")), _ptr(res.get("hess")), _ptr(viol)))
        return res, viol

    def bind(self, what: int, x, y=None, obj_weight: float = 1.0, out=None, wait: bool = True):
        """Validate the buffers once and return a zero-argument callable that issues the same evaluation
        again (solver loops call the same buffers every iteration; this keeps the per-call host cost at
        one ctypes call).  Returns (callable, result dict).  With host arrays, wait=False only enqueues the
        copies and the kernel (ctd_eval_async): call synchronize() before reading the results."""
        x = self._check_in(x, self.nvar, "x")
        y = self._check_in(y, self.ncon, "y")
        dev = _is_torch(x)
        out = dict(out or {})
        res = {}
        for bit, key, n in ((OBJ, "obj", 1), (GRAD, "grad", self.nvar), (CONS, "cons", self.ncon),
                            (JAC, "jac", self.nnzj), (HESS, "hess", self.nnzh)):
            if what & bit:
                buf = out.get(key)
                res[key] = self._check_in(buf, n, key, output=True) if buf is not None else self._empty_like(x, n)
        self._follow_stream(x)   # (resolved once, here: the callable keeps issuing on that stream)
        fn = self._lib.ctd_eval_dev if dev else (self._lib.ctd_eval if wait else self._lib.ctd_eval_async)
        args = (self._h, int(what), _ptr(x), _ptr(y), C.c_double(float(obj_weight)), _ptr(res.get("obj")),
                _ptr(res.get("grad")), _ptr(res.get("cons")), _ptr(res.get("jac")), _ptr(res.get("hess")))
        keep = (x, y, res)

        def call(_fn=fn, _args=args, _keep=keep):
            code = _fn(*_args)
            if code != 0:
                _lib.check(code)

        return call, res

    def bind_with_norms(self, what: int, x, y=None, obj_weight: float = 1.0, out=None, viol=None):
        """`bind` for `evaluate_with_norms` (CUDA tensors): buffers validated once, one ctypes call per
        evaluation afterwards.  Returns (callable, result dict, viol tensor)."""
        import torch
        x = self._check_in(x, self.nvar, "x")
        y = self._check_in(y, self.ncon, "y")
        if not _is_torch(x) or not (what & CONS):
            raise TypeError("bind_with_norms needs CUDA tensors and CONS among the outputs")
        out = dict(out or {})
        res = {}
        for bit, key, n in ((OBJ, "obj", 1), (GRAD, "grad", self.nvar), (CONS, "cons", self.ncon),
                            (JAC, "jac", self.nnzj), (HESS, "hess", self.nnzh)):
            if what & bit:
                buf = out.get(key)
                res[key] = self._check_in(buf, n, key, output=True) if buf is not None else self._empty_like(x, n)
        if viol is None:
            viol = torch.empty(self.batch, dtype=torch.float64, device=x.device)
        self._follow_stream(x)
        fn = self._lib.ctd_eval_norms_dev
        args = (self._h, int(what), _ptr(x), _ptr(y), C.c_double(float(obj_weight)), _ptr(res.get("obj")),
                _ptr(res.get("grad")), _ptr(res.get("cons")), _ptr(res.get("jac")), _ptr(res.get("hess")), _ptr(viol))
        keep = (x, y, res, viol)

        def call(_fn=fn, _args=args, _keep=keep):
            code = _fn(*_args)
            if code != 0:
                _lib.check(code)

        return call, res, viol

    def norms(self, x):
        """Per-instance (objective, max constraint violation): what a batched MPC driver gathers."""
        x = self._check_in(x, self.nvar, "x")
        if _is_torch(x):
            import torch
            obj = torch.empty(self.batch, dtype=torch.float64, device=x.device)
            viol = torch.empty(self.batch, dtype=torch.float64, device=x.device)
            self._follow_stream(x)
            _lib.check(self._lib.ctd_norms_dev(self._h, _ptr(x), _ptr(obj), _ptr(viol)))
        else:
            obj, viol = np.empty(self.batch), np.empty(self.batch)
            _lib.check(self._lib.ctd_norms(self._h, _ptr(x), _ptr(obj), _ptr(viol)))
        return obj, viol

    def violation(self, c, viol=None):
        """max_k max(lcon_k - c_k, c_k - ucon_k, 0) per instance, from constraint values on the device."""
        import torch
        c = self._check_in(c, self.ncon, "c")
        if viol is None:
            viol = torch.empty(self.batch, dtype=torch.float64, device=c.device)
        self._follow_stream(c)
        _lib.check(self._lib.ctd_violation_dev(self._h, _ptr(c), _ptr(viol)))
        return viol

    def unpack(self, x, y=None, want=None):
        """Solution arrays of build_OCP_solution (src/DOCP_data.jl:514-633) from an NLP point x and
        multipliers y: dict with time, state, control, variable (+ costate, path_dual, boundary_dual when y
        is given); arrays are [node, component] ([instance, node, component] for batches)."""
        x = self._check_in(x, self.nvar, "x")
        y = self._check_in(y, self.ncon, "y")
        N, B = self.steps, self.batch
        shapes = {"time": (N + 1,), "state": (N + 1, self.NLP_x), "control": (N * self.control_steps + 1, self.NLP_u),
                  "variable": (self.NLP_v,), "costate": (N, self.NLP_x), "path_dual": (N + 1, self.path_cons),
                  "boundary_dual": (self.boundary_cons,)}
        keys = [k for k in shapes if (want is None or k in want)
                and (y is not None or k in ("time", "state", "control", "variable"))]
        out, sol = {}, _lib.Solution()
        for k in keys:
            shp = shapes[k] if B == 1 else (B,) + shapes[k]
            if _is_torch(x):
                import torch
                out[k] = torch.zeros(shp, dtype=torch.float64, device=x.device)
            else:
                out[k] = np.zeros(shp)
            if int(np.prod(shp)) > 0:
                setattr(sol, k, out[k].data_ptr() if _is_torch(x) else out[k].ctypes.data)
        self._follow_stream(x)
        fn = self._lib.ctd_unpack_dev if _is_torch(x) else self._lib.ctd_unpack
        _lib.check(fn(self._h, _ptr(x), _ptr(y), C.byref(sol)))
        return out

    def control_time(self, time):
        """T_control of build_OCP_solution (src/DOCP_data.jl:555-566) from the `time` array of `unpack`."""
        t = np.ascontiguousarray(time.cpu().numpy() if _is_torch(time) else time, dtype=np.float64)
        out = np.empty(t.shape[:-1] + (self.steps * self.control_steps + 1,))
        _lib.check(self._lib.ctd_control_time(self._h, t.ctypes.data_as(_lib._dp), out.ctypes.data_as(_lib._dp)))
        return out

    def pack(self, state=None, control=None, variable=None, stage_control=None):
        """NLP vector of __initial_guess (src/DOCP_variables.jl:122-145) from sampled data: state [N+1, nx],
        control [N+1, nu], variable [nv] (stagewise schemes: stage_control [N, s, nu]); None keeps the 0.1
        default.  numpy in -> numpy out, CUDA tensors in -> CUDA tensor out."""
        N, B = self.steps, self.batch
        given = [a for a in (state, control, variable, stage_control) if a is not None]
        dev = any(_is_torch(a) for a in given)
        sizes = {"state": (N + 1) * self.NLP_x, "control": (N + 1) * self.NLP_u, "variable": self.NLP_v,
                 "stage_control": N * self.stage * self.NLP_u}
        args = []
        for name, a in (("state", state), ("control", control), ("variable", variable), ("stage_control", stage_control)):
            if a is None or sizes[name] == 0:
                args.append(None)
                continue
            if dev != _is_torch(a):
                raise TypeError("all inputs must be host arrays or all CUDA tensors")
            args.append(self._check_in(a, sizes[name], name))
        if dev:
            import torch
            x = torch.empty(self._shape(self.nvar), dtype=torch.float64, device=given[0].device)
            fn = self._lib.ctd_pack_dev
        else:
            x = np.empty(self._shape(self.nvar))
            fn = self._lib.ctd_pack
        self._follow_stream(*given)
        _lib.check(fn(self._h, *[_ptr(a) for a in args], _ptr(x)))
        return x

    def set_params(self, params):
        par = np.ascontiguousarray(params, dtype=np.float64).ravel()
        _lib.check(self._lib.ctd_set_params(self._h, par.ctypes.data_as(_lib._dp), len(par)))

    def synchronize(self):
        _lib.check(self._lib.ctd_sync(self._h))

    # ---- multi-GPU norm gather through the library's own NCCL communicator (ctd_comm_* of the C ABI) ----
    @staticmethod
    def nccl_unique_id() -> bytes:
        """128 bytes created by ONE rank; hand them to every rank before `comm_init`."""
        lib = _lib.load()
        buf = C.create_string_buffer(128)
        _lib.check(lib.ctd_nccl_unique_id(buf))
        return buf.raw

    def comm_init(self, unique_id: bytes, rank: int, world: int):
        """Join the communicator of the batch (every rank calls this with the same id)."""
        if len(unique_id) != 128:
            raise ValueError("unique_id must be the 128 bytes of nccl_unique_id()")
        _lib.check(self._lib.ctd_comm_init(self._h, C.c_char_p(unique_id), int(rank), int(world)))
        self.comm_world = int(world)

    def gather_norms(self, local, gathered, slot: int = 0):
        """Enqueue the all-gather of `local` (2 x batch CUDA tensor: objective row, violation row) into
        `gathered` (world x 2 x batch) behind everything issued on this handle so far; slot < 4."""
        _lib.check(self._lib.ctd_gather_norms_dev(self._h, _ptr(local), _ptr(gathered), int(slot)))

    def comm_wait(self, slot: int = 0):
        """Make this handle's stream wait until the gather that last used `slot` has landed."""
        _lib.check(self._lib.ctd_comm_wait(self._h, int(slot)))

    # ---- fused norm gather over peer memory (ctd_peer_* of the C ABI) -----------------------------------------
    def peer_init(self):
        """Map every rank's gather buffer into every rank (after `comm_init`; all ranks call it)."""
        _lib.check(self._lib.ctd_peer_init(self._h))
        self._peer_views = {}

    def bind_gather(self, what: int, x, y=None, obj_weight: float = 1.0, out=None, slot: int = 0):
        """`bind_with_norms` whose launch also pushes (objective, violation) of every instance into slot `slot` of
        every rank's gather buffer (ctd_eval_gather_dev).  Returns (callable, result dict)."""
        x = self._check_in(x, self.nvar, "x")
        y = self._check_in(y, self.ncon, "y")
        if not _is_torch(x) or not (what & CONS) or not (what & OBJ):
            raise TypeError("bind_gather needs CUDA tensors and CONS, OBJ among the outputs")
        out = dict(out or {})
        res = {}
        for bit, key, n in ((GRAD, "grad", self.nvar), (CONS, "cons", self.ncon), (JAC, "jac", self.nnzj), (HESS, "hess", self.nnzh)):
            if what & bit:
                buf = out.get(key)
                res[key] = self._check_in(buf, n, key, output=True) if buf is not None else self._empty_like(x, n)
        self._follow_stream(x)
        fn = self._lib.ctd_eval_gather_dev
        args = (self._h, int(what), _ptr(x), _ptr(y), C.c_double(float(obj_weight)), _ptr(res.get("grad")),
                _ptr(res.get("cons")), _ptr(res.get("jac")), _ptr(res.get("hess")), C.c_int32(int(slot)))
        keep = (x, y, res)

        def call(_fn=fn, _args=args, _keep=keep):
            code = _fn(*_args)
            if code != 0:
                _lib.check(code)

        return call, res

    def peer_wait(self, slot: int = 0, overlapped: bool = False):
        """Enqueue the consumer of `slot` (waits for all ranks' norms, copies them out, acknowledges the round) on this
        handle's stream, or with `overlapped` on its communication stream (`comm_wait(slot)` then joins it); returns the
        (world, 2, batch) CUDA tensor view of the slot."""
        import torch
        ptr = C.c_void_p()
        _lib.check(self._lib.ctd_peer_wait(self._h, C.c_int32(int(slot)), C.c_int32(1 if overlapped else 0), C.byref(ptr)))
        view = self._peer_views.get(slot)
        if view is None or view[0] != ptr.value:
            class _Dev:   # a foreign device pointer as a tensor (CUDA array interface)
                pass
            d = _Dev()
            d.__cuda_array_interface__ = {"shape": (self.comm_world, 2, self.batch), "typestr": "<f8",
                                          "data": (int(ptr.value), False), "version": 3, "strides": None}
            view = (ptr.value, torch.as_tensor(d, device=torch.device("cuda", self.device)))
            self._peer_views[slot] = view
        return view[1]

    def set_stream(self, stream=None):
        """Run this handle's work on a caller-owned CUDA stream (a `torch.cuda.Stream`, a raw cudaStream_t as int,
        or None for the handle's own stream): evaluations then follow the caller's stream order and can be
        captured into the caller's CUDA graphs."""
        ptr = getattr(stream, "cuda_stream", stream)
        _lib.check(self._lib.ctd_set_stream(self._h, C.c_void_p(int(ptr)) if ptr else None))
        self._stream_mode, self._stream_set = "own", None   # an explicit choice: stop following torch's current stream

    @property
    def cuda_stream(self) -> int:
        return int(self._lib.ctd_stream(self._h) or 0)

    @property
    def launch_count(self) -> int:
        return int(self._lib.ctd_launch_count(self._h))

    COUNTER_NAMES = ("neval_obj", "neval_grad", "neval_cons", "neval_jac", "neval_hess", "calls", "host_calls",
                     "norm_calls")

    def counters(self) -> dict:
        """The library's per-handle call counters (`ctd_counters`): what `NLPModels.Counters` hold for the
        reference's models (printed by test/benchmark.jl:81-125)."""
        out = np.zeros(len(self.COUNTER_NAMES), dtype=np.int64)
        _lib.check(self._lib.ctd_counters(self._h, out.ctypes.data_as(C.POINTER(C.c_int64))))
        return dict(zip(self.COUNTER_NAMES, (int(v) for v in out)))

    def reset_counters(self) -> None:
        _lib.check(self._lib.ctd_reset_counters(self._h))

    def stats(self) -> dict:
        """Counters, model sizes, kernel launches and the wall time of the blocking host calls (`ctd_stats_json`)."""
        import json
        n = int(self._lib.ctd_stats_json(self._h, None, 0))
        buf = C.create_string_buffer(n)
        self._lib.ctd_stats_json(self._h, buf, n)
        return json.loads(buf.value.decode())

    # ---- the reference's builder seam ---------------------------------------------------------------
    def nlp_model(self, initial_guess=None):
        if self._nlp is None or initial_guess is not None:
            self._nlp = B200NLPModel(self, initial_guess)
        return self._nlp

    def adnlp_model_builder(self, initial_guess=None, backend: str = "b200", **kwargs):
        """Same slot as build_adnlp_model(initial_guess; backend, kwargs...) (src/collocation.jl:90-153)."""
        if backend not in ("b200", "manual"):
            raise ValueError("this backend evaluates the :manual pattern on the GPU; backend must be 'b200'")
        return self.nlp_model(initial_guess)


class B200NLPModel:
    """NLPModels-style model [3P API] whose callbacks run on the B200 (drop-in for the ADNLPModel built at
    src/collocation.jl:137-149)."""

    def __init__(self, docp: DOCP, initial_guess=None):
        self.docp = docp
        lvar, uvar, lcon, ucon = docp.bounds()
        x0 = docp.initial_guess() if initial_guess is None else np.ascontiguousarray(initial_guess, dtype=np.float64)
        self.meta = Meta(docp.nvar, docp.ncon, docp.nnzj, docp.nnzh, x0, lvar, uvar, lcon, ucon,
                         minimize=not docp.max, name=f"{docp.problem}:{docp.scheme}")
        self.counters = {"neval_obj": 0, "neval_grad": 0, "neval_cons": 0, "neval_jac": 0, "neval_hess": 0}

    def obj(self, x):
        self.counters["neval_obj"] += 1
        f = self.docp.evaluate(OBJ, x)["obj"]
        if _is_torch(f):
            return f if self.docp.batch > 1 else f.reshape(())
        return f if self.docp.batch > 1 else float(f[0])

    def grad(self, x, g=None):
        self.counters["neval_grad"] += 1
        return self.docp.evaluate(GRAD | OBJ, x, out={"grad": g})["grad"]

    def cons(self, x, c=None):
        self.counters["neval_cons"] += 1
        return self.docp.evaluate(CONS, x, out={"cons": c})["cons"]

    def jac_structure(self):
        return self.docp.jacobian_structure()

    def jac_coord(self, x, vals=None):
        self.counters["neval_jac"] += 1
        return self.docp.evaluate(JAC, x, out={"jac": vals})["jac"]

    def hess_structure(self):
        return self.docp.hessian_structure()

    def hess_coord(self, x, y, vals=None, obj_weight: float = 1.0):
        self.counters["neval_hess"] += 1
        return self.docp.evaluate(HESS, x, y, obj_weight, out={"hess": vals})["hess"]

    def eval_all(self, x, y, obj_weight: float = 1.0, out=None):
        """obj, grad, cons, jac_coord and hess_coord in one launch."""
        for k in self.counters:
            self.counters[k] += 1
        return self.docp.evaluate(CONS | JAC | HESS | GRAD | OBJ, x, y, obj_weight, out=out)


# ---- reference entry points ------------------------------------------------------------------------
def discretize(ocp, discretizer: Optional[Collocation] = None, **kwargs) -> DOCP:
    """CTDirect.discretize(ocp, Collocation(...)) (src/CTDirect.jl:31-50)."""
    d = discretizer or Collocation()
    if isinstance(d, DirectShooting):
        return DOCP(ocp, grid_size=d.grid_size, scheme=d.scheme, control_steps=d.control_steps, **kwargs)
    return DOCP(ocp, grid_size=d.grid_size, scheme=d.scheme, time_grid=d.time_grid, **kwargs)


def direct_transcription(ocp, grid_size: int = 250, disc_method: str = "midpoint", time_grid=None, init=None, **kwargs):
    """Legacy spelling (test/tmp/percival.jl:24): returns (docp, nlp)."""
    docp = DOCP(ocp, grid_size=grid_size, scheme=disc_method, time_grid=time_grid, **kwargs)
    return docp, docp.nlp_model(init)


def nlp_model(docp: DOCP):
    return docp.nlp_model()


def DOCP_objective(x, docp: DOCP):       # test/benchmark.jl:112
    return docp.nlp_model().obj(x)


def DOCP_constraints(c, x, docp: DOCP):  # DOCP_constraints! test/benchmark.jl:125
    return docp.nlp_model().cons(x, c)


def DOCP_initial_guess(docp: DOCP):      # test/benchmark.jl:81-86
    return docp.initial_guess()

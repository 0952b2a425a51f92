"""Symbolic optimal control problems consumed by the kernel generator.

Each OCP is given through the callables CTDirect obtains from CTModels:
dynamics(t, x, u, v), lagrange(t, x, u, v), mayer(x0, xf, v), path(t, x, u, v), boundary(x0, xf, v),
box triplets (lb, index, ub) and the default initial guess.  `p` holds runtime parameters (one row per
instance in batched mode).  Sources in the reference tree:
  goddard / goddard_all      test/problems/goddard.jl:7-49, 87-158
  double_integrator_*        test/problems/double_integrator.jl:4-19, 40-56, 79-99 (free t0 and tf)
  beam                       test/problems/beam.jl:3-17
  stagewise_scalar           test/ci/test_discretization_stagewise.jl:1-14
  bolza_freetf               test/problems/bolza.jl:3-18
  vanderpol                  test/problems/vanderpol.jl:3-19
  estimate_*                 test/problems/autonomous_system.jl:6-90 (no control: NLP_u = 0)
  robot                      NOT in the reference (synthetic, after the COPS robot arm)
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Callable, List, Optional, Tuple

import sympy as sp

INF = float("inf")

# A structural zero: `u * ZERO` is numerically 0 but keeps its dependency on u for sparsity DETECTION
# (the reference's @def Goddard writes the dynamics as F0(x) + u*F1(x) with F1 = [0, Tmax/m, -b*Tmax],
# test/problems/goddard.jl:7-15,32, and the tracer of the `:optimized` backend does not look at values).
ZERO = sp.Symbol("CTD_STRUCT_ZERO")


@dataclass
class OCP:
    name: str
    nx: int
    nu: int
    nv: int
    params: List[Tuple[str, float]]
    t0: tuple  # ("const", value) | ("var", index in v, 0-based) | ("param", index)
    tf: tuple
    dynamics: Callable
    lagrange: Optional[Callable] = None
    mayer: Optional[Callable] = None
    criterion: str = "min"
    path: Optional[Callable] = None
    path_lb: List[float] = field(default_factory=list)
    path_ub: List[float] = field(default_factory=list)
    boundary: Optional[Callable] = None
    bc_lb: List[float] = field(default_factory=list)
    bc_ub: List[float] = field(default_factory=list)
    state_box: List[Tuple[float, int, float]] = field(default_factory=list)   # (lb, 1-based idx, ub)
    control_box: List[Tuple[float, int, float]] = field(default_factory=list)
    variable_box: List[Tuple[float, int, float]] = field(default_factory=list)
    init_state: Optional[List[float]] = None
    init_control: Optional[List[float]] = None
    init_variable: Optional[List[float]] = None
    synthetic: bool = False

    @property
    def np(self):
        return len(self.path_lb)

    @property
    def nb(self):
        return len(self.bc_lb)

    @property
    def npar(self):
        return len(self.params)

    @property
    def free_t0(self):
        return self.t0[0] == "var"

    @property
    def free_tf(self):
        return self.tf[0] == "var"

    def time_expr(self, which, v, p):
        kind, val = self.t0 if which == 0 else self.tf
        if kind == "const":
            return sp.Float(val) if val != int(val) else sp.Integer(int(val))
        if kind == "var":
            return v[val]
        return p[val]


def _goddard_dyn(t, x, u, v, p):
    Cd, beta, b, Tmax = p
    r, vel, m = x
    D = Cd * vel**2 * sp.exp(-beta * (r - 1))
    return [vel, -D / m - 1 / r**2 + u[0] * Tmax / m, -b * Tmax * u[0]]


def _goddard_def_dyn(t, x, u, v, p):
    """F0(x) + u*F1(x) of the @def version: the first component keeps a structural dependency on u."""
    f = _goddard_dyn(t, x, u, v, p)
    return [f[0] + u[0] * ZERO, f[1], f[2]]


_GODDARD_PARAMS = [("Cd", 310.0), ("beta", 500.0), ("b", 2.0), ("Tmax", 3.5)]

goddard = OCP(
    name="goddard", nx=3, nu=1, nv=1, params=_GODDARD_PARAMS, t0=("const", 0.0), tf=("var", 0),
    dynamics=_goddard_def_dyn,
    mayer=lambda x0, xf, v, p: xf[0], criterion="max",
    boundary=lambda x0, xf, v, p: [x0[0], x0[1], x0[2], xf[2]],
    bc_lb=[1.0, 0.0, 1.0, 0.6], bc_ub=[1.0, 0.0, 1.0, 0.6],
    state_box=[(1.0, 1, 1.1), (0.0, 2, 0.1), (0.6, 3, 1.0)],
    control_box=[(0.0, 1, 1.0)], variable_box=[(0.01, 1, INF)],
    init_state=[1.01, 0.05, 0.8],
)

goddard_all = OCP(
    name="goddard_all", nx=3, nu=1, nv=1, params=_GODDARD_PARAMS, t0=("const", 0.0), tf=("var", 0),
    dynamics=_goddard_dyn,
    mayer=lambda x0, xf, v, p: xf[0], criterion="max",
    path=lambda t, x, u, v, p: [x[1], u[0], x[0] + x[1] + x[2] + u[0] + v[0]],
    path_lb=[-INF, -INF, 0.0], path_ub=[0.1, 1.0, INF],
    boundary=lambda x0, xf, v, p: [x0[0], x0[1], x0[2], xf[2]],
    bc_lb=[1.0, 0.0, 1.0, 0.6], bc_ub=[1.0, 0.0, 1.0, 0.6],
    state_box=[(1.0, 1, INF), (0.0, 2, INF), (0.0, 3, 1.0)],
    control_box=[(0.0, 1, INF)], variable_box=[(0.01, 1, INF)],
    init_state=[1.01, 0.05, 0.8],
)

double_integrator_minenergy = OCP(
    name="double_integrator_minenergy", nx=2, nu=1, nv=0, params=[("T", 2.0)],
    t0=("const", 0.0), tf=("param", 0),
    dynamics=lambda t, x, u, v, p: [x[1], u[0]],
    lagrange=lambda t, x, u, v, p: u[0] ** 2,
    boundary=lambda x0, xf, v, p: [x0[0], x0[1], xf[0], xf[1]],
    bc_lb=[0.0, 0.0, 1.0, 0.0], bc_ub=[0.0, 0.0, 1.0, 0.0],
)

double_integrator_mintf = OCP(
    name="double_integrator_mintf", nx=2, nu=1, nv=1, params=[], t0=("const", 0.0), tf=("var", 0),
    dynamics=lambda t, x, u, v, p: [x[1], u[0]],
    mayer=lambda x0, xf, v, p: v[0],
    boundary=lambda x0, xf, v, p: [x0[0], x0[1], xf[0], xf[1]],
    bc_lb=[0.0, 0.0, 1.0, 0.0], bc_ub=[0.0, 0.0, 1.0, 0.0],
    control_box=[(-1.0, 1, 1.0)], variable_box=[(0.05, 1, INF)],
)

# free initial AND final time, t0 -> max (test/problems/double_integrator.jl:79-99; swept over the six schemes at
# grid_size 50 by test/ci/test_discretization.jl:74-80).  `0.01 <= tf - t0` is a function of the variable alone:
# CTModels keeps it with the boundary constraints (functions of (x0, xf, v)), after the two conditions on x.
double_integrator_freet0tf = OCP(
    name="double_integrator_freet0tf", nx=2, nu=1, nv=2, params=[], t0=("var", 0), tf=("var", 1),
    dynamics=lambda t, x, u, v, p: [x[1], u[0]],
    mayer=lambda x0, xf, v, p: v[0], criterion="max",
    boundary=lambda x0, xf, v, p: [x0[0], x0[1], xf[0], xf[1], v[1] - v[0]],
    bc_lb=[0.0, 0.0, 1.0, 0.0, 0.01], bc_ub=[0.0, 0.0, 1.0, 0.0, INF],
    control_box=[(-1.0, 1, 1.0)], variable_box=[(0.05, 1, 10.0), (0.05, 2, 10.0)],
)

# beam (from bocop), Lagrange cost with a state box (test/problems/beam.jl:3-17; swept over the six schemes by
# test/ci/test_discretization.jl:66-72)
beam = OCP(
    name="beam", nx=2, nu=1, nv=0, params=[], t0=("const", 0.0), tf=("const", 1.0),
    dynamics=lambda t, x, u, v, p: [x[1], u[0]],
    lagrange=lambda t, x, u, v, p: u[0] ** 2,
    boundary=lambda x0, xf, v, p: [x0[0], x0[1], xf[0], xf[1]],
    bc_lb=[0.0, 1.0, 0.0, -1.0], bc_ub=[0.0, 1.0, 0.0, -1.0],
    state_box=[(0.0, 1, 0.1)], control_box=[(-10.0, 1, 10.0)],
)

stagewise_scalar = OCP(
    name="stagewise_scalar", nx=1, nu=1, nv=0, params=[], t0=("const", 0.0), tf=("const", 1.0),
    dynamics=lambda t, x, u, v, p: [u[0]],
    lagrange=lambda t, x, u, v, p: u[0] ** 2,
    boundary=lambda x0, xf, v, p: [x0[0], xf[0]],
    bc_lb=[0.0, 1.0], bc_ub=[0.0, 1.0],
    control_box=[(0.0, 1, 2.0)],
)

bolza_freetf = OCP(
    name="bolza_freetf", nx=1, nu=1, nv=1, params=[], t0=("const", 0.0), tf=("var", 0),
    dynamics=lambda t, x, u, v, p: [v[0] * u[0]],
    lagrange=lambda t, x, u, v, p: sp.Rational(1, 2) * u[0] ** 2,
    mayer=lambda x0, xf, v, p: v[0],
    boundary=lambda x0, xf, v, p: [x0[0], xf[0]],
    bc_lb=[0.0, 1.0], bc_ub=[0.0, 1.0],
    state_box=[(0.0, 1, INF)], variable_box=[(0.1, 1, INF)],
)

vanderpol = OCP(
    name="vanderpol", nx=2, nu=1, nv=0, params=[("omega", 1.0), ("epsilon", 1.0)],
    t0=("const", 0.0), tf=("const", 2.0),
    dynamics=lambda t, x, u, v, p: [x[1], p[1] * p[0] * (1 - x[0] ** 2) * x[1] - p[0] ** 2 * x[0] + u[0]],
    lagrange=lambda t, x, u, v, p: sp.Rational(1, 2) * (x[0] ** 2 + x[1] ** 2 + u[0] ** 2),
    boundary=lambda x0, xf, v, p: [x0[0], x0[1]],
    bc_lb=[1.0, 0.0], bc_ub=[1.0, 0.0],
)

_PI = 3.14159265358979323846


def _robot_dyn(t, x, u, v, p):
    L = p[0]
    Iphi = ((L - x[0]) ** 3 + x[0] ** 3) / 3
    Itheta = Iphi * sp.sin(x[4]) ** 2
    return [x[1], u[0] / L, x[3], u[1] / Itheta, x[5], u[2] / Iphi]


robot = OCP(
    name="robot", nx=6, nu=3, nv=1, params=[("L", 5.0)], t0=("const", 0.0), tf=("var", 0),
    dynamics=_robot_dyn,
    mayer=lambda x0, xf, v, p: v[0],
    boundary=lambda x0, xf, v, p: list(x0) + list(xf),
    bc_lb=[4.5, 0, 0, 0, _PI / 4, 0, 4.5, 0, 2 * _PI / 3, 0, _PI / 4, 0],
    bc_ub=[4.5, 0, 0, 0, _PI / 4, 0, 4.5, 0, 2 * _PI / 3, 0, _PI / 4, 0],
    state_box=[(0.0, 1, 5.0), (-_PI, 3, _PI), (0.0, 5, _PI)],
    control_box=[(-1.0, 1, 1.0), (-1.0, 2, 1.0), (-1.0, 3, 1.0)],
    variable_box=[(0.1, 1, INF)],
    init_state=[4.5, 0.0, 1.0, 0.0, _PI / 4, 0.0], init_control=[0.0, 0.0, 0.0], init_variable=[1.0],
    synthetic=True,
)

# parameter estimation without control (NLP_u = 0): test/problems/autonomous_system.jl:6-90
estimate_initial_condition = OCP(
    name="estimate_initial_condition", nx=2, nu=0, nv=2, params=[], t0=("const", 0.0), tf=("const", _PI / 2),
    dynamics=lambda t, x, u, v, p: [-x[1], x[0]],
    mayer=lambda x0, xf, v, p: (xf[0] - 0) ** 2 + (xf[1] - 1) ** 2,
    boundary=lambda x0, xf, v, p: [x0[0] - v[0], x0[1] - v[1]],
    bc_lb=[0.0, 0.0], bc_ub=[0.0, 0.0],
)

estimate_rotation_rate = OCP(
    name="estimate_rotation_rate", nx=2, nu=0, nv=1, params=[], t0=("const", 0.0), tf=("const", 1.0),
    dynamics=lambda t, x, u, v, p: [v[0] * (-x[1]), v[0] * x[0]],
    mayer=lambda x0, xf, v, p: (xf[0] - 0) ** 2 + (xf[1] - 1) ** 2 + sp.Float(0.01) * v[0] ** 2,
    boundary=lambda x0, xf, v, p: [x0[0] - 1, x0[1] - 0],
    bc_lb=[0.0, 0.0], bc_ub=[0.0, 0.0],
    variable_box=[(0.0, 1, 10.0)],
)

ALL = {o.name: o for o in (
    goddard, goddard_all, double_integrator_minenergy, double_integrator_mintf, double_integrator_freet0tf, beam,
    stagewise_scalar,
    bolza_freetf, vanderpol, robot, estimate_initial_condition, estimate_rotation_rate,
)}

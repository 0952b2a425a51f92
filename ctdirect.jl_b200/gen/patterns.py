"""Sparsity patterns of the discretised OCP, build-time mirror used by the kernel generator.

Follows the reference's `:manual` patterns: `DOCP_Jacobian_pattern` / `DOCP_Hessian_pattern`
(src/ode/trapeze.jl:149-303, midpoint.jl:163-300, irk.jl:315-496, irk_stagewise.jl:468-638) built
from `add_nonzero_block!` (src/ode/common.jl:295-312) and `SparseArrays.sparse` (CSC: sorted by
column, then row; duplicates merged).  The Hessian is reduced to its lower triangle, as ADNLPModels
does [3P].  All indices are 1-based like the reference's.

The generator only uses this on tiny probe grids to derive the per-step value offsets that are baked
into the CUDA kernels; the library builds the full structure itself (csrc/pattern_host.cpp).
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

SCHEMES = {
    # name: (stage, irk, stagewise, final_control)
    "trapeze": (0, False, False, True),
    "midpoint": (0, False, False, False),
    "gauss_legendre_1_constant_control": (1, True, False, False),
    "gauss_legendre_2_constant_control": (2, True, False, False),
    "gauss_legendre_3_constant_control": (3, True, False, False),
    "gauss_legendre_2": (2, True, True, False),
    "gauss_legendre_3": (3, True, True, False),
    "euler": (0, False, False, False),            # explicit Euler, src/ode/euler.jl
    "euler_implicit": (0, False, False, False),   # implicit Euler, src/ode/euler.jl (explicit=false)
    # midpoint rule with several controls per time step (the DirectShooting discretizer: src/direct_shooting.jl:55-71,
    # midpoint.jl:20,57-65,98-113,145-154); the number of control steps is a compile-time constant of the element
    "midpoint_cs2": (0, False, False, False),
    "midpoint_cs3": (0, False, False, False),
    "midpoint_cs10": (0, False, False, False),
}

CONTROL_STEPS = {"midpoint_cs2": 2, "midpoint_cs3": 3, "midpoint_cs10": 10}


def is_midpoint(scheme: str) -> bool:
    return scheme == "midpoint" or scheme in CONTROL_STEPS


@dataclass
class Dims:
    nx: int
    nu: int
    nv: int
    np: int
    nb: int
    lagrange: bool
    mayer: bool


@dataclass
class Layout:
    """Scheme dimensions: trapeze.jl:14-42, midpoint.jl:17-39, irk.jl:138-160, irk_stagewise.jl:136-163."""
    scheme: str
    dims: Dims
    N: int

    def __post_init__(self):
        d = self.dims
        self.stage, self.irk, self.stagewise, self.final_control = SCHEMES[self.scheme]
        s = self.stage
        self.control_steps = CONTROL_STEPS.get(self.scheme, 1)
        self.control_block = d.nu * s if self.stagewise else d.nu * self.control_steps   # midpoint.jl:20
        self.step_block = d.nx + self.control_block + (s * d.nx if self.irk else 0)
        self.eqs_block = d.nx * (1 + s) if self.irk else d.nx
        self.cblock = self.eqs_block + d.np
        self.nvar = self.N * self.step_block + d.nx + d.nv + (d.nu if self.final_control else 0)
        self.ncon = self.N * self.cblock + d.np + d.nb
        # variables local to one step: own block, then what it touches in the next block
        self.n_own = self.step_block
        self.n_next = d.nx + (d.nu if self.scheme == "trapeze" else 0)
        self.nz = self.n_own + self.n_next + d.nv


class _Triplets:
    def __init__(self):
        self.i = []
        self.j = []

    def block(self, i0, i1, j0, j1, sym=False):
        if i1 < i0 or j1 < j0:
            return  # empty Julia range
        ii, jj = np.meshgrid(np.arange(i0, i1 + 1), np.arange(j0, j1 + 1), indexing="ij")
        self.i.append(ii.ravel())
        self.j.append(jj.ravel())
        if sym:
            self.i.append(jj.ravel())
            self.j.append(ii.ravel())

    def one(self, i, j):
        self.i.append(np.array([i]))
        self.j.append(np.array([j]))

    def csc(self, ncols, lower=False):
        i = np.concatenate(self.i).astype(np.int64)
        j = np.concatenate(self.j).astype(np.int64)
        if lower:
            keep = i >= j
            i, j = i[keep], j[keep]
        key = np.unique(j * (2**32) + i)
        cols = key >> 32
        rows = key & (2**32 - 1)
        colptr = np.zeros(ncols + 1, dtype=np.int64)
        np.add.at(colptr, cols, 1)
        colptr = np.cumsum(colptr)
        return rows, cols, colptr


def jacobian_pattern(L: Layout, pattern_mode: str = "reference"):
    d, N, B, cb = L.dims, L.N, L.step_block, L.cblock
    nx, nu, np_, nb, s = d.nx, d.nu, d.np, d.nb, L.stage
    T = _Triplets()
    v0, v1 = L.nvar - d.nv + 1, L.nvar
    for i in range(1, N + 1):
        co = (i - 1) * cb
        vo = (i - 1) * B
        dyn0, dyn1 = co + 1, co + nx
        if is_midpoint(L.scheme):
            T.block(dyn0, dyn1, vo + 1, vo + B + nx)
            T.block(co + nx + 1, co + cb, vo + 1, vo + nx + nu)
            T.block(co + 1, co + cb, v0, v1)
        elif L.scheme in ("euler", "euler_implicit"):
            # euler.jl:166-262 (the Lagrange-cost lines repeat entries of the dynamics block)
            p0, p1 = co + nx + 1, co + cb
            T.block(dyn0, dyn1, vo + 1, vo + nx)
            T.block(dyn0, dyn1, vo + nx + 1, vo + B + nx)
            T.block(dyn0, dyn1, v0, v1)
            T.block(p0, p1, vo + 1, vo + nx)
            if L.scheme == "euler_implicit" and pattern_mode == "corrected" and i >= 2:
                # node i of the implicit rule reads the control of block i-1 (euler.jl:66-69: u(]t_i, t_i+1]) = U_i,
                # u(t_1) = U_1); the reference's :manual pattern lists block i (euler.jl:214-216)
                T.block(p0, p1, vo - B + nx + 1, vo - B + nx + nu)
            else:
                T.block(p0, p1, vo + nx + 1, vo + nx + nu)
            T.block(p0, p1, v0, v1)
        elif L.scheme == "trapeze":
            p0, p1 = co + nx + 1, co + cb
            T.block(dyn0, dyn1, vo + 1, vo + nx)
            T.block(dyn0, dyn1, vo + nx + 1, vo + nx + nu + nx)
            T.block(dyn0, dyn1, vo + 2 * nx + nu + 1, vo + 2 * nx + 2 * nu)
            T.block(p0, p1, vo + 1, vo + nx)
            T.block(p0, p1, vo + nx + 1, vo + nx + nu)
            T.block(p0, p1, v0, v1)  # trapeze.jl:203 (quirk Q1: path rows only)
            if pattern_mode == "corrected":
                T.block(dyn0, dyn1, v0, v1)
        else:
            st0, st1 = co + nx + 1, co + (s + 1) * nx
            p0, p1 = co + (s + 1) * nx + 1, co + cb
            xi0, xi1 = vo + 1, vo + nx
            ui1 = vo + nx + L.control_block
            ki0, ki1 = vo + nx + L.control_block + 1, vo + B
            T.block(dyn0, dyn1, xi0, xi1)
            T.block(dyn0, dyn1, ki0, vo + B + nx)
            T.block(dyn0, dyn1, v0, v1)
            T.block(st0, st1, xi0, ki1)
            T.block(st0, st1, v0, v1)
            T.block(p0, p1, xi0, ui1)
            T.block(p0, p1, v0, v1)
    co = N * cb
    vo = N * B
    xf0, xf1 = vo + 1, vo + nx
    if L.scheme == "trapeze":
        uf0, uf1 = vo + nx + 1, vo + nx + nu
    else:
        uf0, uf1 = vo - B + nx + 1, vo - B + nx + L.control_block
    T.block(co + 1, co + np_, xf0, xf1)
    T.block(co + 1, co + np_, uf0, uf1)
    T.block(co + 1, co + np_, v0, v1)
    bo = co + np_
    T.block(bo + 1, bo + nb, 1, nx)
    T.block(bo + 1, bo + nb, xf0, xf1)
    T.block(bo + 1, bo + nb, v0, v1)
    if (L.stagewise or L.scheme in ("euler", "euler_implicit")) and d.lagrange and pattern_mode == "reference":
        T.one(L.ncon, nx)  # irk_stagewise.jl:550-552, euler.jl:257-259 (quirk Q3)
    return T.csc(L.nvar)


def hessian_pattern(L: Layout, pattern_mode: str = "reference"):
    d, N, B = L.dims, L.N, L.step_block
    nx, nu = d.nx, d.nu
    T = _Triplets()
    v0, v1 = L.nvar - d.nv + 1, L.nvar
    T.block(v0, v1, v0, v1)
    for i in range(1, N + 1):
        vo = (i - 1) * B
        if is_midpoint(L.scheme):
            T.block(vo + 1, vo + B + nx, vo + 1, vo + B + nx)
            T.block(vo + 1, vo + B + nx, v0, v1, sym=True)
        elif L.scheme in ("euler", "euler_implicit"):
            # euler.jl:269-355
            x0_, x1_ = vo + 1, vo + nx
            xn0, xn1 = vo + B + 1, vo + B + nx
            u0_, u1_ = vo + nx + 1, vo + nx + nu
            T.block(u0_, u1_, u0_, u1_)
            T.block(u0_, u1_, v0, v1, sym=True)
            if L.scheme == "euler_implicit":
                T.block(xn0, xn1, xn0, xn1)
                T.block(xn0, xn1, u0_, u1_, sym=True)
                T.block(xn0, xn1, v0, v1, sym=True)
            T.block(x0_, x1_, x0_, x1_)
            T.block(x0_, x1_, u0_, u1_, sym=True)
            T.block(x0_, x1_, v0, v1, sym=True)
        elif L.scheme == "trapeze":
            T.block(vo + 1, vo + 2 * B, vo + 1, vo + 2 * B)
            T.block(vo + 1, vo + 2 * B, v0, v1, sym=True)
        else:
            T.block(vo + 1, vo + B, vo + 1, vo + B)
            T.block(vo + 1, vo + B, v0, v1, sym=True)
    vo = N * B
    xf0, xf1 = vo + 1, vo + nx
    if L.irk:
        uf0, uf1 = vo - B + nx + 1, vo - B + nx + L.control_block
        T.block(xf0, xf1, xf0, xf1)
        T.block(uf0, uf1, uf0, uf1)
        T.block(xf0, xf1, uf0, uf1, sym=True)
        if L.stagewise:
            T.block(xf0, uf1, v0, v1, sym=True)  # irk_stagewise.jl:624: empty range (quirk Q2)
            if pattern_mode == "corrected":
                T.block(xf0, xf1, v0, v1, sym=True)
        else:
            T.block(xf0, xf1, v0, v1, sym=True)
        T.block(uf0, uf1, v0, v1, sym=True)
        T.block(1, nx, xf0, xf1, sym=True)
    elif is_midpoint(L.scheme) or L.scheme in ("euler", "euler_implicit"):
        T.block(1, nx, xf0, xf1, sym=True)  # quirk Q4: always (midpoint.jl:286-291, euler.jl:345-350)
    else:
        if d.mayer or d.nb > 0:
            T.block(1, nx, xf0, xf1, sym=True)
    return T.csc(L.nvar, lower=True)


def find_pos(rows, colptr, r, c):
    """0-based position of 1-based (r, c) in the CSC arrays, or -1."""
    a, b = colptr[c - 1], colptr[c]
    k = a + np.searchsorted(rows[a:b], r)
    if k < b and rows[k] == r:
        return int(k)
    return -1


def optimized_patterns(L: Layout, step_J, step_H, extra_J, extra_H, shift_from=None):
    """`:optimized` (detected) patterns assembled from per-element structure: `step_J` = [(row r, local a)],
    `step_H` = [(local a, local b)] of one time step; `extra_*` = global 1-based (i, j) entries of the
    final-path and boundary(+Mayer) elements.  Returns the same CSC triples as the manual builders."""
    d, N, B, cb = L.dims, L.N, L.step_block, L.cblock
    nloc = L.n_own + L.n_next

    def gcol(i, a):
        return (i - 1) * B + a + 1 if a < nloc else L.nvar - d.nv + (a - nloc) + 1

    # shift_from: local rows >= shift_from of step i are rows of the NEXT step's block (implicit Euler with path
    # constraints, see gen/elements.py); those of step N do not exist (the final-path element holds node N+1)
    TJ, TH = _Triplets(), _Triplets()
    for i in range(1, N + 1):
        for (r, a) in step_J:
            if shift_from is not None and r >= shift_from:
                if i < N:
                    TJ.one(i * cb + r + 1, gcol(i, a))
                continue
            TJ.one((i - 1) * cb + r + 1, gcol(i, a))
        for (a, b) in step_H:
            TH.one(gcol(i, a), gcol(i, b))
            TH.one(gcol(i, b), gcol(i, a))
    for (i, j) in extra_J:
        TJ.one(i, j)
    for (i, j) in extra_H:
        TH.one(i, j)
        TH.one(j, i)
    if not TH.i:                       # no second-order term at all: empty pattern
        hz = (np.zeros(0, dtype=np.int64), np.zeros(0, dtype=np.int64), np.zeros(L.nvar + 1, dtype=np.int64))
    else:
        hz = TH.csc(L.nvar, lower=True)
    return TJ.csc(L.nvar), hz

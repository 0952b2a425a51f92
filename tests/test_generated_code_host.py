"""The generated element code (csrc/generated/*.cuh), compiled for the HOST with g++ and assembled by a
small Python loop, against the CPU oracle.  This checks the generator's analytic derivatives and slot tables
without a GPU; the CUDA tiling / staging logic is covered by the `-m gpu` tests."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from conftest import assert_close, column_scale, hess_scale, synthetic_point

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GEN = os.path.join(ROOT, "ctdirect.jl_b200", "csrc", "generated")
_dp = C.POINTER(C.c_double)


def _build(name, tmp):
    so = os.path.join(tmp, f"lib{name}.so")
    cmd = ["g++", "-O1", "-std=c++17", "-shared", "-fPIC", "-w", "-D__device__=", "-D__host__=", "-D__forceinline__=inline",
           "-D__restrict__=", f"-DCTD_ELEMENT_HEADER=\"{os.path.join(GEN, name + '.cuh')}\"", f"-DCTD_ELEMENT={name}",
           os.path.join(ROOT, "tests", "host_harness", "harness.cpp"), "-o", so]
    subprocess.check_call(cmd)
    lib = C.CDLL(so)
    lib.elem_step.restype = C.c_double
    lib.elem_bc.restype = C.c_double
    return lib


def _p(a):
    return a.ctypes.data_as(_dp)


def _assemble(lib, m, x, y, sigma, sgrid, path_shift=False):
    """obj, grad, cons, jac, hess of oracle model m's problem from the host-compiled element code.
    `path_shift` (implicit Euler with path constraints): the local rows EQS.. of step i are the path rows of node
    i+1 -- rows of the NEXT step's block (gen/elements.py); node 1 is evaluated with the final-path code on block 1,
    node N+1 by the final-path element."""
    d = np.zeros(30, dtype=np.int32)
    lib.elem_dims(d.ctypes.data_as(C.POINTER(C.c_int)))
    (NZ, CB, NJO, NJN, NJV, NHO, NHN, NVV, NGO, NGN, NV, FPN, NP, NFPJ, NFPH, BCN, NB, NBCJ, NBCH, NBCG, NOWN, NNEXT,
     BLK, NX, NPAR, T0K, TFK, T0I, TFI, EQS) = (int(v) for v in d)
    consts = np.zeros(2)
    lib.elem_consts(_p(consts))

    def table(i):
        buf = np.zeros(3 * 4096, dtype=np.int32)
        n = lib.elem_table(i, buf.ctypes.data_as(C.POINTER(C.c_int)))
        return buf[:3 * n].reshape(n, 3)
    JO, JN, JV, HO, HN, FPJ, FPH, BCJ, BCH, GO, GN, GV, BCG = (table(i) for i in range(13))
    N, nvar, ncon = m.N, m.nvar, m.ncon
    jr, jc = m.jac_structure()
    hr, hc = m.hess_structure()
    jpos = {(int(r), int(c)): k for k, (r, c) in enumerate(zip(jr, jc))}
    hpos = {(int(r), int(c)): k for k, (r, c) in enumerate(zip(hr, hc))}
    cons, jac, hess, grad = np.zeros(ncon), np.zeros(m.nnzj), np.zeros(m.nnzh), np.zeros(nvar)
    par = np.array(m._params if hasattr(m, "_params") else [], dtype=float)
    v = x[nvar - NV:]
    p = np.zeros(max(NPAR, 1))
    p[:len(par)] = par
    t0 = consts[0] if T0K == 0 else (v[T0I] if T0K == 1 else p[T0I])
    tf = consts[1] if TFK == 0 else (v[TFI] if TFK == 1 else p[TFI])

    def targs(i):   # step i (1-based): ti, tp, hh and seeds
        si, sp = sgrid[i - 1], (sgrid[i] if i <= N else sgrid[i - 1])
        ti, tp = t0 + si * (tf - t0), t0 + sp * (tf - t0)
        t = np.zeros(3 + 3 * max(NV, 1))
        t[0], t[1], t[2] = ti, tp, tp - ti
        for k in range(NV):
            di = (si if (TFK == 1 and k == TFI) else 0.0) + ((1 - si) if (T0K == 1 and k == T0I) else 0.0)
            dq = (sp if (TFK == 1 and k == TFI) else 0.0) + ((1 - sp) if (T0K == 1 and k == T0I) else 0.0)
            t[3 + k], t[3 + NV + k], t[3 + 2 * NV + k] = di, dq, dq - di
        return t

    def gcol(i, a):
        return (i - 1) * BLK + a + 1 if a < NOWN + NNEXT else nvar - NV + (a - NOWN - NNEXT) + 1

    def addh(ga, gb, val):
        k = hpos.get((max(ga, gb), min(ga, gb)))
        if k is not None:
            hess[k] += val

    def grow(i, r):   # 1-based global row of local row r of step i (None: no such row)
        if path_shift and r >= EQS:
            return i * CB + r + 1 if i < N else None
        return (i - 1) * CB + r + 1

    obj = 0.0
    for i in range(1, N + 1):
        z = np.concatenate([x[(i - 1) * BLK:(i - 1) * BLK + NOWN + NNEXT], v])
        lam = y[(i - 1) * CB:i * CB].copy()
        if path_shift:
            lam[EQS:] = y[i * CB + EQS:(i + 1) * CB] if i < N else 0.0
        o = {k: np.zeros(max(n, 1)) for k, n in (("c", CB), ("jo", NJO), ("jn", NJN), ("jv", NJV), ("ho", NHO),
                                                 ("hn", NHN), ("hv", NVV), ("go", NGO), ("gn", NGN), ("gv", NV))}
        obj += lib.elem_step(_p(z), _p(targs(i)), _p(p), _p(lam), C.c_double(sigma), *[_p(o[k]) for k in
                             ("c", "jo", "jn", "jv", "ho", "hn", "hv", "go", "gn", "gv")])
        for r in range(CB):
            if grow(i, r) is not None:
                cons[grow(i, r) - 1] = o["c"][r]
        for tab, vals in ((JO, o["jo"]), (JN, o["jn"])):
            for n, (r, a, _) in enumerate(tab):
                k = jpos.get((grow(i, r), gcol(i, a)))
                if k is not None:
                    jac[k] = vals[n]
        for n, (r, kv, _) in enumerate(JV):
            k = jpos.get((grow(i, r), nvar - NV + kv + 1))
            if k is not None:
                jac[k] = o["jv"][n]
        for tab, vals in ((HO, o["ho"]), (HN, o["hn"])):
            for n, (a, b, _) in enumerate(tab):
                addh(gcol(i, a), gcol(i, b), vals[n])
        for a in range(NV):
            for b in range(a + 1):
                addh(nvar - NV + a + 1, nvar - NV + b + 1, o["hv"][a * (a + 1) // 2 + b])
        for n, (a, _, _) in enumerate(GO):
            grad[gcol(i, a) - 1] += o["go"][n]
        for n, (a, _, _) in enumerate(GN):
            grad[gcol(i, a) - 1] += o["gn"][n]
        grad[nvar - NV:] += o["gv"][:NV]
    final_control = m.nvar == N * BLK + NX + NV + m.nu     # trapeze: a control at the final node
    if NP > 0:
        cb_ = FPN - NX - NV
        z = np.concatenate([x[N * BLK:N * BLK + NX + m.nu], v]) if final_control else \
            np.concatenate([x[(N - 1) * BLK + NX:(N - 1) * BLK + NX + cb_], x[N * BLK:N * BLK + NX], v])

        def fcol(a):
            if a >= FPN - NV:
                return nvar - NV + (a - (FPN - NV)) + 1
            if final_control:
                return N * BLK + a + 1
            return (N - 1) * BLK + NX + a + 1 if a < cb_ else N * BLK + (a - cb_) + 1
        lam = y[N * CB:N * CB + NP].copy()
        c_, j_, h_ = np.zeros(NP), np.zeros(max(NFPJ, 1)), np.zeros(max(NFPH, 1))
        lib.elem_fpath(_p(z), _p(targs(N + 1)), _p(p), _p(lam), _p(c_), _p(j_), _p(h_))
        cons[N * CB:N * CB + NP] = c_
        for n, (r, a, _) in enumerate(FPJ):
            k = jpos.get((N * CB + r + 1, fcol(a)))
            if k is not None:
                jac[k] = j_[n]
        for n, (a, b, _) in enumerate(FPH):
            addh(fcol(a), fcol(b), h_[n])
    if NP > 0 and path_shift:   # path rows of node 1: the final-path code on (U_1 block, x_1, v)
        cb_ = FPN - NX - NV
        z = np.concatenate([x[NX:NX + cb_], x[:NX], v])
        icol = lambda a: nvar - NV + (a - (FPN - NV)) + 1 if a >= FPN - NV else (NX + a + 1 if a < cb_ else (a - cb_) + 1)
        lam = y[EQS:EQS + NP].copy()
        c_, j_, h_ = np.zeros(NP), np.zeros(max(NFPJ, 1)), np.zeros(max(NFPH, 1))
        lib.elem_fpath(_p(z), _p(targs(1)), _p(p), _p(lam), _p(c_), _p(j_), _p(h_))
        cons[EQS:EQS + NP] = c_
        for n, (r, a, _) in enumerate(FPJ):
            k = jpos.get((EQS + r + 1, icol(a)))
            if k is not None:
                jac[k] = j_[n]
        for n, (a, b, _) in enumerate(FPH):
            addh(icol(a), icol(b), h_[n])
    z = np.concatenate([x[:NX], x[N * BLK:N * BLK + NX], v])
    bcol = lambda a: a + 1 if a < NX else (N * BLK + (a - NX) + 1 if a < 2 * NX else nvar - NV + (a - 2 * NX) + 1)
    lam = y[ncon - NB:].copy() if NB else np.zeros(1)
    c_, j_, h_, g_ = np.zeros(max(NB, 1)), np.zeros(max(NBCJ, 1)), np.zeros(max(NBCH, 1)), np.zeros(max(NBCG, 1))
    obj += lib.elem_bc(_p(z), _p(p), _p(lam), C.c_double(sigma), _p(c_), _p(j_), _p(h_), _p(g_))
    cons[ncon - NB:] = c_[:NB]
    for n, (r, a, _) in enumerate(BCJ):
        k = jpos.get((ncon - NB + r + 1, bcol(a)))
        if k is not None:
            jac[k] = j_[n]
    for n, (a, b, _) in enumerate(BCH):
        addh(bcol(a), bcol(b), h_[n])
    for n, (a, _, _) in enumerate(BCG):
        grad[bcol(a) - 1] += g_[n]
    return obj, grad, cons, jac, hess


CASES = [("goddard", "midpoint", "reference"), ("goddard", "midpoint", "optimized"), ("goddard_all", "trapeze", "reference"),
         ("goddard_all", "trapeze", "corrected"), ("goddard", "gauss_legendre_3", "reference"),
         ("goddard_all", "gauss_legendre_2_constant_control", "reference"), ("robot", "trapeze", "optimized"),
         ("double_integrator_minenergy", "gauss_legendre_2", "reference"), ("bolza_freetf", "euler_implicit", "reference"),
         ("goddard_all", "euler", "reference"), ("vanderpol", "gauss_legendre_3_constant_control", "reference"),
         ("estimate_rotation_rate", "midpoint", "reference"),
         # the reference's own scheme sweep (test/ci/test_discretization.jl:62-80): free t0 AND tf / Lagrange + state box
         ("double_integrator_freet0tf", "trapeze", "reference"), ("double_integrator_freet0tf", "trapeze", "corrected"),
         ("double_integrator_freet0tf", "midpoint", "reference"), ("double_integrator_freet0tf", "euler", "reference"),
         ("double_integrator_freet0tf", "euler_implicit", "reference"), ("double_integrator_freet0tf", "gauss_legendre_2", "reference"),
         ("double_integrator_freet0tf", "gauss_legendre_3", "reference"), ("beam", "trapeze", "reference"),
         ("beam", "gauss_legendre_3", "reference"), ("beam", "euler_implicit", "reference"),
         # implicit Euler WITH path constraints: node i reads the control of block i-1 (euler.jl:66-69); corrected pattern only
         ("goddard_all", "euler_implicit", "corrected"), ("goddard_all", "euler_implicit", "optimized"),
         # several controls per step (DirectShooting)
         ("goddard_all", "midpoint_cs2", "optimized"), ("double_integrator_minenergy", "midpoint_cs10", "optimized")]


if os.environ.get("CTD_HOST_ALL"):   # opt-in: every compiled (problem, scheme, pattern mode) triple, ~15 min on 8 cores
    CASES = [tuple(l.strip()[2:].split("__")) for l in open(os.path.join(GEN, "MANIFEST.txt")) if l.strip()]
    CASES = [(c[0], c[1], c[2] if len(c) > 2 else "reference") for c in CASES]


@pytest.mark.parametrize("problem,scheme,mode", CASES)
def test_host_compiled_elements_match_oracle(oracle_mod, tmp_path, problem, scheme, mode):
    name = f"E_{problem}__{scheme}" + ("" if mode == "reference" else "__" + mode)
    lib = _build(name, str(tmp_path))
    for N, grid in ((1, None), (2, None), (5, None), (6, np.array([0.0, 0.1, 0.15, 0.4, 0.6, 0.9, 1.0]))):
        m = oracle_mod.OracleModel(problem, scheme, N, time_grid=grid, pattern_mode=mode)
        m._params = {"goddard": [310.0, 500.0, 2.0, 3.5], "goddard_all": [310.0, 500.0, 2.0, 3.5], "robot": [5.0],
                     "double_integrator_minenergy": [2.0], "vanderpol": [1.0, 1.0]}.get(problem, [])
        x, y = synthetic_point(m, seed=11 + N)
        sgrid = np.arange(N + 1) / N if grid is None else grid
        obj, grad, cons, jac, hess = _assemble(lib, m, x, y, 0.7, sgrid, path_shift=(scheme == "euler_implicit" and m.np_ > 0))
        g_ref, j_ref, h_ref = m.derivs(x, y, 0.7)
        tag = f"{name}/N={N}"
        assert_close(cons, m.cons(x), scale=max(1.0, np.abs(x).max()), what=tag + " cons")
        assert_close(jac, j_ref, scale=column_scale(m.jac_structure(), j_ref, m.nvar), what=tag + " jac")
        assert_close(hess, h_ref, scale=hess_scale(m, x, y, 0.7, h_ref), what=tag + " hess")
        assert_close(grad, g_ref, scale=1e-3 * max(1.0, np.abs(g_ref).max()), what=tag + " grad")
        assert_close(obj, m.obj(x), scale=1.0, what=tag + " obj")

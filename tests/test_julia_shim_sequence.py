"""The Julia glue (ctdirect.jl_b200/julia/CTDirectB200.jl) cannot be executed in this image (no Julia).  Its call
sequence into the C ABI is replayed by a C program compiled against include/ctdirect_b200.h
(tests/host_harness/julia_shim_sequence.c): create with the normalized grid, dims, info, bounds, x0, structures, the five
NLPModels callbacks with host pointers, the fused call, the solution unpacking -- for every scheme name the shim's
type -> name table (`scheme_name`) can produce.  The results are compared with the oracle."""
import os
import re
import subprocess

import numpy as np
import pytest

from conftest import assert_close, column_scale, hess_scale, synthetic_point

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "host_harness", "julia_shim_sequence.c")
SHIM = os.path.join(ROOT, "ctdirect.jl_b200", "julia", "CTDirectB200.jl")

# reference discretization struct (src/ode/*.jl) -> the scheme symbol that selects it (src/DOCP_data.jl:307-336)
REFERENCE_TYPES = {
    "Trapeze": "trapeze", "Midpoint": "midpoint", "Euler(explicit)": "euler", "Euler(implicit)": "euler_implicit",
    "Gauss_Legendre_1": "gauss_legendre_1_constant_control", "Gauss_Legendre_2": "gauss_legendre_2_constant_control",
    "Gauss_Legendre_3": "gauss_legendre_3_constant_control", "Gauss_Legendre_2_Stagewise": "gauss_legendre_2",
    "Gauss_Legendre_3_Stagewise": "gauss_legendre_3",
}


def _build(tmp):
    import ctdirect_b200
    exe = os.path.join(tmp, "julia_shim_sequence")
    libdir = os.path.dirname(ctdirect_b200.LIB_PATH)
    subprocess.run(["gcc", "-std=c99", "-O1", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), SRC, "-o", exe,
                    "-L", libdir, "-lctdirect_b200", f"-Wl,-rpath,{libdir}", "-lm"], check=True)
    return exe


def test_shim_scheme_table_matches_the_reference_switch():
    """`scheme_name` in the Julia file maps every discretization struct of the reference to the name the library knows
    (the round-1 shim lower-cased the type name, which turns constant-control IRK into stagewise and loses implicit Euler)."""
    src = open(SHIM).read()
    table = dict(re.findall(r'T == "(\w+)" && return "(\w+)"', src))
    for typ, name in REFERENCE_TYPES.items():
        if typ.startswith("Euler"):
            continue
        assert table[typ] == name, (typ, table.get(typ))
    assert re.search(r'T == "Euler" && return getfield\(disc, :_explicit\) \? "euler" : "euler_implicit"', src)
    assert "normalized_grid=docp.time.normalized_grid" in src and "fixed_grid" not in src.split("function b200_model_builder")[1].split("return function")[1]
    for fn in ("direct_transcription", "nlp_model", "solve", "B200DOCPModelGPU", "eval_report"):
        assert re.search(rf"\b{fn}\b", src), fn


def test_harness_compiles_against_the_header(ctd, tmp_path):
    """The header is plain C (the harness is built with gcc -std=c99 -Werror) and the library exports what it uses."""
    exe = _build(str(tmp_path))
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 1 and "usage" in r.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("scheme", sorted(set(REFERENCE_TYPES.values())))
def test_shim_call_sequence_on_gpu(ctd, oracle_mod, tmp_path, scheme):
    exe = _build(str(tmp_path))
    cases = [("goddard", 60, 0), ("double_integrator_freet0tf", 50, 0)]   # free tf / free t0 and tf: where fixed_grid is garbage
    if scheme in ("trapeze", "gauss_legendre_2", "gauss_legendre_3"):
        cases.append(("goddard", 60, 1))
    if scheme == "midpoint":   # the DirectShooting discretizer: docp.time.control_steps = 2, detected pattern
        cases.append(("goddard", 60, 2))
    for problem, N, mode in cases:
        extra = ["2"] if mode == 2 else []
        o = oracle_mod.OracleModel(problem, "midpoint_cs2" if mode == 2 else scheme, N,
                                   pattern_mode=("reference", "corrected", "optimized")[mode])
        x, y = synthetic_point(o, seed=5)
        if problem == "double_integrator_freet0tf":
            x[-2:] = (0.3, 2.1)
        fx, fy, fo = (str(tmp_path / n) for n in ("x.bin", "y.bin", "out.bin"))
        x.tofile(fx); y.tofile(fy)
        r = subprocess.run([exe, problem, scheme, str(N), str(mode), fx, fy, fo] + extra, capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        raw = np.fromfile(fo, dtype=np.float64)
        hdr = raw[:8].view(np.int64)
        nvar, ncon, nnzj, nnzh, is_max, n_, nx, nu = (int(v) for v in hdr)
        assert (nvar, ncon, nnzj, nnzh) == (o.nvar, o.ncon, o.nnzj, o.nnzh) and n_ == N and bool(is_max) == o.is_max
        pos = [8]

        def take(n, dtype=np.float64):
            a = raw[pos[0]:pos[0] + n]
            pos[0] += n
            return a.view(dtype)
        lvar, uvar, lcon, ucon, x0 = take(nvar), take(nvar), take(ncon), take(ncon), take(nvar)
        jr, jc, hr, hc = (take(n, np.int64) for n in (nnzj, nnzj, nnzh, nnzh))
        f, g, c, jv, hv = take(1)[0], take(nvar), take(ncon), take(nnzj), take(nnzh)
        tm, st = take(N + 1), take((N + 1) * nx)
        fused = (take(1)[0], take(nvar), take(ncon), take(nnzj), take(nnzh))
        for a, b in zip((lvar, uvar, lcon, ucon), o.bounds()):
            assert np.array_equal(a, b)
        assert np.array_equal(x0, o.x0())
        for a, b in zip((jr, jc, hr, hc), o.jac_structure() + o.hess_structure()):
            assert np.array_equal(a, b)
        g_ref, j_ref, h_ref = o.derivs(x, y, 0.7)
        sh = hess_scale(o, x, y, 0.7, h_ref)
        for kind, (f_, g_, c_, jv_, hv_) in (("callbacks", (f, g, c, jv, hv)), ("fused", fused)):
            tag = f"shim sequence ({kind}) {problem}/{scheme}/mode{mode}"
            assert_close(c_, o.cons(x), scale=max(1.0, np.abs(x).max()), what=tag + " cons")
            assert_close(jv_, j_ref, scale=column_scale((jr, jc), j_ref, nvar), what=tag + " jac")
            assert_close(hv_, h_ref, scale=sh, what=tag + " hess")
            assert_close(g_, g_ref, scale=1e-3 * max(1.0, np.abs(g_ref).max()), what=tag + " grad")
            assert_close(f_, o.obj(x), scale=1.0, what=tag + " obj")
        u = o.unpack(x, y)
        assert np.array_equal(tm, u["time"]) and np.array_equal(st.reshape(N + 1, nx), u["state"])

"""The kernel generator: committed sources are up to date, value-slot tables are consistent."""
import numpy as np
import pytest


def test_generated_sources_are_up_to_date():
    from ctdirect_b200.gen import generate
    assert generate.main(["--check"]) == 0


@pytest.mark.parametrize("problem,scheme", [("goddard", "midpoint"), ("goddard_all", "trapeze"),
                                             ("double_integrator_minenergy", "gauss_legendre_2"),
                                             ("robot", "gauss_legendre_2_constant_control")])
def test_slot_tables_cover_interior_blocks(problem, scheme):
    from ctdirect_b200.gen import elements as el, ocps
    E = el.build_step_element(ocps.ALL[problem], scheme)
    # every slot of an interior Jacobian column block is written exactly once (own + next)
    slots = sorted([e[2] for e in E.j_own] + [e[2] for e in E.j_next])
    assert slots == list(range(E.j_block))
    assert sorted(e[2] for e in E.h_own) == list(range(E.h_block))
    # closed forms of SURVEY.md §8a for midpoint
    if scheme == "midpoint":
        d = E.L.dims
        cb = d.nx + d.np
        assert E.j_block == d.nx * (d.nx + cb) + d.nu * cb


def test_generated_derivatives_match_sympy_numerically():
    """Spot check of one generated expression set against direct symbolic differentiation at 30 digits."""
    import sympy as sp
    from ctdirect_b200.gen import elements as el, ocps
    E = el.build_step_element(ocps.goddard, "midpoint")
    rng = np.random.default_rng(0)
    vals = {s: sp.Float(float(v), 30) for s, v in zip(E.z, [1.01, 0.05, 0.8, 0.4, 1.011, 0.051, 0.79, 0.2])}
    vals.update({s: sp.Float(float(v), 30) for s, v in zip(E.lam, rng.standard_normal(3))})
    vals.update({s: sp.Float(v, 30) for s, v in zip(E.p, [310.0, 500.0, 2.0, 3.5])})
    si, sp1 = sp.Rational(3, 10), sp.Rational(4, 10)
    tf = vals[E.z[7]]
    vals.update({E.ts.ti: si * tf, E.ts.tp: sp1 * tf, E.ts.h: (sp1 - si) * tf, E.sg: sp.Float(0.7, 30),
                 E.ts.dti[0]: si, E.ts.dtp[0]: sp1, E.ts.dh[0]: sp1 - si})
    inter = {s: e for s, e in E.inter}
    # direct formulation: substitute the intermediates and times as functions of z, differentiate with sympy
    z = E.z
    full = lambda e: sp.sympify(e).subs(inter).subs({E.ts.ti: si * z[7], E.ts.tp: sp1 * z[7], E.ts.h: (sp1 - si) * z[7]})
    lagr = sum(E.lam[r] * full(E.rows[r]) for r in range(3))
    num = lambda e: float(sp.sympify(e).subs(inter).subs(vals).subs(vals).evalf(25))
    for (a, b, slot, expr) in E.h_own + E.h_next:
        ref = float(sp.diff(lagr, z[a], z[b]).subs(vals).evalf(25))
        got = num(expr)
        assert abs(got - ref) <= 1e-12 * max(1.0, abs(ref)), (a, b, got, ref)
    for (r, a, slot, expr) in E.j_own + E.j_next + E.j_v:
        ref = float(sp.diff(full(E.rows[r]), z[a]).subs(vals).evalf(25))
        assert abs(num(expr) - ref) <= 1e-12 * max(1.0, abs(ref)), (r, a)

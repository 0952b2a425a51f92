"""Ad-hoc parity report on a GPU box: max relative error per output for a few (problem, scheme, N)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import ctdirect_b200 as ctd
from oracle.oracle import OracleModel
from conftest import synthetic_point

def rel(a, b, floor):
    a = np.asarray(a, float); b = np.asarray(b, float)
    s = np.maximum(np.maximum(np.abs(a), np.abs(b)), floor)
    e = np.abs(a - b) / s
    k = int(np.argmax(e)) if e.size else 0
    return (float(e.max()) if e.size else 0.0), k

cases = [(p, s, m, N) for (p, s, m) in ctd.list_models() for N in (1, 2, 3, 130, 300)]
if len(sys.argv) > 1:
    cases = [c for c in cases if sys.argv[1] in c[0] + "/" + c[1]]
worst = {}
nbad = 0
for (p, s, m, N) in cases:
    o = OracleModel(p, s, N, pattern_mode=m)
    d = ctd.DOCP(p, N, s, pattern_mode=m)
    nlp = d.nlp_model()
    x, y = synthetic_point(o, seed=7 + N)
    g_ref, j_ref, h_ref = o.derivs(x, y, 0.7)
    c_ref, f_ref = o.cons(x), o.obj(x)
    r = nlp.eval_all(x, y, obj_weight=0.7)
    sep = {"cons": nlp.cons(x), "jac": nlp.jac_coord(x), "hess": nlp.hess_coord(x, y, obj_weight=0.7),
           "grad": nlp.grad(x), "obj": np.atleast_1d(nlp.obj(x))}
    ref = {"cons": c_ref, "jac": j_ref, "hess": h_ref, "grad": g_ref, "obj": np.atleast_1d(f_ref)}
    line = []
    for k in ("cons", "jac", "hess", "grad", "obj"):
        floor = max(1e-300, 1e-3 * float(np.abs(ref[k]).max())) if k != "cons" else max(1.0, float(np.abs(x).max()))
        e1, k1 = rel(r[k], ref[k], floor)
        e2, k2 = rel(sep[k], ref[k], floor)
        line.append(f"{k}:{max(e1, e2):.1e}")
        if max(e1, e2) > 1e-12 or not np.isfinite(max(e1, e2)):
            nbad += 1
            kk = k1 if e1 >= e2 else k2
            print("  BAD", p, s, m, N, k, "fused" if e1 >= e2 else "separate", "idx", kk, r[k].flat[kk] if e1 >= e2 else sep[k].flat[kk], ref[k].flat[kk])
        worst[k] = max(worst.get(k, 0.0), e1, e2)
    print(p, s, m, N, " ".join(line), "dropped", o.dropped, flush=True)
print("WORST", worst, "nbad", nbad)

"""Experiment: one fused evaluation (cons+jac+hess) as TWO concurrent kernels -- Jacobian side (cons + jac) on one stream,
Hessian on another -- with a fork / join per evaluation, so that consecutive evaluations stay serially dependent.
Both variants are captured into CUDA graphs of ITER evaluations on rotating buffer sets and replayed, so the host launch
rate does not enter.   python scripts/split_eval_check.py [problem scheme N]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import ctdirect_b200 as ctd

problem = sys.argv[1] if len(sys.argv) > 1 else "goddard"
scheme = sys.argv[2] if len(sys.argv) > 2 else "midpoint"
N = int(sys.argv[3]) if len(sys.argv) > 3 else 100000
dev = torch.device("cuda", 0)
dA = ctd.DOCP(problem, N, scheme, stream="own")     # fused, and the cons + jac half
dB = ctd.DOCP(problem, N, scheme, stream="own")     # the Hessian half
sA = torch.cuda.ExternalStream(dA.cuda_stream, device=dev)
sB = torch.cuda.ExternalStream(dB.cuda_stream, device=dev)
rng = np.random.default_rng(0)
NSETS, ITER = 6, 12
sets = []
for s in range(NSETS):
    x = torch.from_numpy(dA.initial_guess() * (1 + 1e-3 * rng.uniform(-1, 1, dA.nvar))).to(dev)
    y = torch.from_numpy(rng.standard_normal(dA.ncon)).to(dev)
    out = {"cons": torch.empty(dA.ncon, dtype=torch.float64, device=dev), "jac": torch.empty(dA.nnzj, dtype=torch.float64, device=dev),
           "hess": torch.empty(dA.nnzh, dtype=torch.float64, device=dev)}
    sets.append((x, y, out))
fused = [dA.bind(ctd.CONS | ctd.JAC | ctd.HESS, x, y, 1.0, out=o)[0] for x, y, o in sets]
# cons and jac are separate compiled output sets (1, 2); the split variant launches cons+jac as two kernels on stream A
cons = [dA.bind(ctd.CONS, x, y, 1.0, out={"cons": o["cons"]})[0] for x, y, o in sets]
jac = [dA.bind(ctd.JAC, x, y, 1.0, out={"jac": o["jac"]})[0] for x, y, o in sets]
hess = [dB.bind(ctd.HESS, x, y, 1.0, out={"hess": o["hess"]})[0] for x, y, o in sets]
for k in range(NSETS):
    fused[k](); cons[k](); jac[k](); hess[k]()
torch.cuda.synchronize()
ref = {k: v.clone() for k, v in sets[0][2].items()}


def capture(body):
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=sA, capture_error_mode="thread_local"):
        body()
    torch.cuda.synchronize()
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    return g


def time_graph(g, reps=100):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(sA):
        a.record(sA)
        for _ in range(reps):
            g.replay()
        b.record(sA)
    torch.cuda.synchronize()
    return a.elapsed_time(b) * 1e3 / (reps * ITER)


def body_fused():
    for i in range(ITER):
        fused[i % NSETS]()


def body_split():
    for i in range(ITER):
        k = i % NSETS
        e = torch.cuda.Event(); e.record(sA); sB.wait_event(e)        # fork: the Hessian kernel follows evaluation i-1
        hess[k]()
        cons[k](); jac[k]()
        j = torch.cuda.Event(); j.record(sB); sA.wait_event(j)        # join: evaluation i+1 follows both halves


per = 8 * (dA.nvar + 2 * dA.ncon + dA.nnzj + dA.nnzh)
for name, body in (("fused, one stream", body_fused), ("cons+jac | hess on two streams, fork/join per evaluation", body_split)):
    g = capture(body)
    us = min(time_graph(g) for _ in range(3))
    print(f"{problem}/{scheme}/N={N}  {name:60s} {us:7.2f} us/eval  {per / us / 1e3:7.1f} GB/s  {per / us / 1e3 / 6552.6 * 100:5.1f}% of measured HBM")
for k in ("cons", "jac", "hess"):
    assert torch.equal(sets[0][2][k], ref[k]), k
print("results identical")

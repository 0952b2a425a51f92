"""Experiment: is the multi-stream throughput limited by the host launch rate?  Compares the Python issue
loop with a CUDA graph holding one round of evaluations on NS forked streams."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import ctdirect_b200 as ctd
NS = int(sys.argv[1]) if len(sys.argv) > 1 else 3
ROUNDS = int(sys.argv[2]) if len(sys.argv) > 2 else 2
K = 600
dev = torch.device("cuda", 0)
docps = [ctd.DOCP("goddard", 100000, "midpoint", stream="own") for _ in range(NS)]
d0 = docps[0]
streams = [torch.cuda.ExternalStream(d.cuda_stream, device=dev) for d in docps]
rng = np.random.default_rng(0)
nsets = NS * ROUNDS
sets = []
for s in range(nsets):
    x = torch.from_numpy(d0.initial_guess() * (1 + 1e-3 * rng.uniform(-1, 1, d0.nvar))).to(dev)
    y = torch.from_numpy(rng.standard_normal(d0.ncon)).to(dev)
    out = {"cons": torch.empty(d0.ncon, dtype=torch.float64, device=dev), "jac": torch.empty(d0.nnzj, dtype=torch.float64, device=dev),
           "hess": torch.empty(d0.nnzh, dtype=torch.float64, device=dev)}
    sets.append((x, y, out))
F = ctd.CONS | ctd.JAC | ctd.HESS
calls = [docps[k % NS].bind(F, s[0], s[1], 1.0, out=s[2])[0] for k, s in enumerate(sets)]
for c in calls: c()
torch.cuda.synchronize()

def loop_time():
    e0 = [torch.cuda.Event(enable_timing=True) for _ in streams]; e1 = [torch.cuda.Event(enable_timing=True) for _ in streams]
    for e, st in zip(e0, streams): e.record(st)
    for i in range(K): calls[i % nsets]()
    for e, st in zip(e1, streams): e.record(st)
    torch.cuda.synchronize()
    return max(a.elapsed_time(b) for a, b in zip(e0, e1)) * 1e3 / K

print(f"python loop, {NS} streams: {loop_time():.2f} us/eval  {loop_time():.2f}")
import time
t0 = time.perf_counter()
for i in range(K): calls[i % nsets]()
t1 = time.perf_counter()
torch.cuda.synchronize()
print(f"host issue time per call: {(t1 - t0) / K * 1e6:.2f} us")

g = torch.cuda.CUDAGraph()
main = streams[0]
with torch.cuda.graph(g, stream=main, capture_error_mode="thread_local"):
    fork = torch.cuda.Event(); fork.record(main)
    for st in streams[1:]: st.wait_event(fork)
    for c in calls: c()
    for st in streams[1:]:
        e = torch.cuda.Event(); e.record(st); main.wait_event(e)
torch.cuda.synchronize()
for _ in range(3): g.replay()
torch.cuda.synchronize()
R = K // nsets
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
with torch.cuda.stream(main):
    a.record(main)
    for _ in range(R): g.replay()
    b.record(main)
torch.cuda.synchronize()
print(f"graph of {nsets} evaluations on {NS} streams: {a.elapsed_time(b) * 1e3 / (R * nsets):.2f} us/eval")

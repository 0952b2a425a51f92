"""Throughput of back-to-back fused evaluations: plain stream loop vs CUDA graph replay vs 2 streams."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import ctdirect_b200 as ctd
N = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
F = ctd.CONS | ctd.JAC | ctd.HESS
dev = torch.device("cuda", 0)
rng = np.random.default_rng(0)
def make(nsets):
    d = ctd.DOCP("goddard", N, "midpoint", stream="own")
    x0 = d.initial_guess()
    calls = []
    keep = []
    for s in range(nsets):
        x = torch.from_numpy(x0 * (1 + 1e-3 * rng.uniform(-1, 1, d.nvar))).to(dev)
        y = torch.from_numpy(rng.standard_normal(d.ncon)).to(dev)
        c, res = d.bind(F, x, y, 1.0)
        calls.append(c); keep.append((x, y, res))
    return d, calls, keep
K = 2000
d, calls, keep = make(4)
st = torch.cuda.ExternalStream(d.cuda_stream, device=dev)
def timeit(fn, reps, streams):
    torch.cuda.synchronize()
    e0 = [torch.cuda.Event(enable_timing=True) for _ in streams]; e1 = [torch.cuda.Event(enable_timing=True) for _ in streams]
    for s, e in zip(streams, e0): e.record(s)
    fn(reps)
    for s, e in zip(streams, e1): e.record(s)
    torch.cuda.synchronize()
    return max(a.elapsed_time(b) for a, b in zip(e0, e1)) * 1e3
with torch.cuda.stream(st):
    for i in range(20): calls[i % 4]()
    t = timeit(lambda n: [calls[i % 4]() for i in range(n)], K, [st])
print(f"plain loop        : {t / K:7.2f} us/eval")
# CUDA graph of one rotation (4 launches)
g = torch.cuda.CUDAGraph()
with torch.cuda.stream(st):
    torch.cuda.synchronize()
    g.capture_begin()
    for i in range(4): calls[i]()
    g.capture_end()
    for _ in range(5): g.replay()
    t = timeit(lambda n: [g.replay() for _ in range(n // 4)], K, [st])
print(f"graph (4 launches): {t / K:7.2f} us/eval")
g2 = torch.cuda.CUDAGraph()
with torch.cuda.stream(st):
    torch.cuda.synchronize()
    g2.capture_begin()
    for i in range(40): calls[i % 4]()
    g2.capture_end()
    for _ in range(3): g2.replay()
    t = timeit(lambda n: [g2.replay() for _ in range(n // 40)], K, [st])
print(f"graph (40 launches): {t / K:7.2f} us/eval")
# two handles, two streams, alternating
d2, calls2, keep2 = make(4)
st2 = torch.cuda.ExternalStream(d2.cuda_stream, device=dev)
for i in range(8): calls2[i % 4]()
torch.cuda.synchronize()
def two(n):
    for i in range(n // 2):
        calls[i % 4](); calls2[i % 4]()
t = timeit(two, K, [st, st2])
print(f"2 streams         : {t / K:7.2f} us/eval")
d3, calls3, keep3 = make(4)
st3 = torch.cuda.ExternalStream(d3.cuda_stream, device=dev)
for i in range(8): calls3[i % 4]()
torch.cuda.synchronize()
def three(n):
    for i in range(n // 3):
        calls[i % 4](); calls2[i % 4](); calls3[i % 4]()
t = timeit(three, K - K % 3, [st, st2, st3])
print(f"3 streams         : {t / (K - K % 3):7.2f} us/eval")

"""Per-flagset kernel timing (CUDA events, rotating buffers): python scripts/ktime.py [problem scheme N [batch]]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import ctdirect_b200 as ctd

problem = sys.argv[1] if len(sys.argv) > 1 else "goddard"
scheme = sys.argv[2] if len(sys.argv) > 2 else "midpoint"
N = int(sys.argv[3]) if len(sys.argv) > 3 else 100000
batch = int(sys.argv[4]) if len(sys.argv) > 4 else 1
mode = sys.argv[5] if len(sys.argv) > 5 else "reference"
d = ctd.DOCP(problem, N, scheme, batch=batch, pattern_mode=mode, stream="own")
dev = torch.device("cuda", 0)
rng = np.random.default_rng(0)
x0 = d.initial_guess()
per = 8 * (d.nvar * 2 + 2 * d.ncon + d.nnzj + d.nnzh) * batch
nsets = max(3, int(np.ceil(1.5 * 126e6 / per))) if per < 300e6 else 1
NS_ = int(os.environ.get('KTIME_STREAMS', '1'))
nsets = int(os.environ.get('KTIME_SETS', nsets))
nsets = ((max(nsets, NS_) + NS_ - 1) // NS_) * NS_
sets = []
for s in range(nsets):
    x = torch.from_numpy(np.ascontiguousarray(x0[None, :] * (1 + 1e-3 * rng.uniform(-1, 1, (batch, d.nvar))))).to(dev)
    y = torch.from_numpy(rng.standard_normal((batch, d.ncon))).to(dev)
    out = {"cons": torch.empty((batch, d.ncon), dtype=torch.float64, device=dev),
           "jac": torch.empty((batch, d.nnzj), dtype=torch.float64, device=dev),
           "hess": torch.empty((batch, d.nnzh), dtype=torch.float64, device=dev),
           "grad": torch.empty((batch, d.nvar), dtype=torch.float64, device=dev),
           "obj": torch.empty(batch, dtype=torch.float64, device=dev)}
    sets.append((x, y, out))
stream = torch.cuda.ExternalStream(d.cuda_stream, device=dev)
NS = int(os.environ.get("KTIME_STREAMS", "1"))           # independent evaluations in flight
extra = [ctd.DOCP(problem, N, scheme, batch=batch, pattern_mode=mode, stream="own") for _ in range(NS - 1)]
handles = [d] + extra
streams = [torch.cuda.ExternalStream(h.cuda_stream, device=dev) for h in handles]
# reference floor: a device-to-device copy moving the same number of bytes as the fused evaluation
src = torch.empty(per // 16, dtype=torch.float64, device=dev); dst = [torch.empty_like(src) for _ in range(max(nsets, 2))]
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for i in range(5): dst[i % len(dst)].copy_(src)
e0.record()
for i in range(100): dst[i % len(dst)].copy_(src)
e1.record(); torch.cuda.synchronize()
us = e0.elapsed_time(e1) * 10
print(f"torch copy_ moving {per/1e6:.1f} MB (read+write): {us:.2f} us  {per/us/1e3:.1f} GB/s", flush=True)
names = {1: "cons", 2: "jac", 4: "hess", 16: "obj", 17: "cons+obj", 24: "grad+obj", 7: "cons+jac+hess", 31: "all"}
peak = 6552.6
FLAGSETS = tuple(int(f) for f in os.environ.get('KTIME_FLAGS', '1,2,4,16,24,7,31').split(','))
for F in FLAGSETS:
    calls = [handles[k % NS].bind(F, x, y, 1.0, out={k2: v for k2, v in out.items()})[0] for k, (x, y, out) in enumerate(sets)]
    nb = 8 * batch * ((d.nvar) + (d.ncon if F & 4 else 0) + (d.ncon if F & 1 else 0) + (d.nnzj if F & 2 else 0)
                      + (d.nnzh if F & 4 else 0) + (d.nvar if F & 8 else 0))
    K = int(os.environ.get('KTIME_K', 300 if batch == 1 else 20))
    with torch.cuda.stream(stream):
        for i in range(10):
            calls[i % nsets]()
        torch.cuda.synchronize()
        ev0 = [torch.cuda.Event(enable_timing=True) for _ in streams]
        ev1 = [torch.cuda.Event(enable_timing=True) for _ in streams]
        for e, st in zip(ev0, streams): e.record(st)
        for i in range(K):
            calls[i % nsets]()
        for e, st in zip(ev1, streams): e.record(st)
        torch.cuda.synchronize()
    us = max(a.elapsed_time(b) for a, b in zip(ev0, ev1)) * 1e3 / K
    print(f"{problem}/{scheme}/{mode}/N={N}/B={batch}/streams={NS} F={F:2d} {names[F]:14s} {us:9.2f} us  {nb/1e6:8.2f} MB  {nb/us/1e3:8.1f} GB/s  {nb/us/1e3/peak*100:5.1f}% of measured HBM", flush=True)

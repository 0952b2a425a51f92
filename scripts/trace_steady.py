"""Debug: per-warp timeline of a launch in the steady state of back-to-back launches on one stream (needs a
-DCTD_TRACE build: CTD_LIB=.../libctd_trace.so).  usage: trace_steady.py FLAGS WPC [problem scheme N]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
F = int(sys.argv[1]); WPC = int(sys.argv[2])
AT = 40
os.environ["CTD_TRACE_FILE"] = "/tmp/trace_pair.bin"
os.environ["CTD_TRACE_AT"] = str(AT)
import ctdirect_b200 as ctd
PROB = sys.argv[3] if len(sys.argv) > 3 else "goddard"
SCH = sys.argv[4] if len(sys.argv) > 4 else "midpoint"
NN = int(sys.argv[5]) if len(sys.argv) > 5 else 100000
d = ctd.DOCP(PROB, NN, SCH, stream="own")
dev = torch.device("cuda", 0)
rng = np.random.default_rng(0)
nsets = 6
sets = []
for s in range(nsets):
    x = torch.from_numpy(d.initial_guess() * (1 + 1e-3 * rng.uniform(-1, 1, d.nvar))).to(dev)
    y = torch.from_numpy(rng.standard_normal(d.ncon)).to(dev)
    out = {"cons": torch.empty(d.ncon, dtype=torch.float64, device=dev), "jac": torch.empty(d.nnzj, dtype=torch.float64, device=dev),
           "hess": torch.empty(d.nnzh, dtype=torch.float64, device=dev), "grad": torch.empty(d.nvar, dtype=torch.float64, device=dev),
           "obj": torch.empty(1, dtype=torch.float64, device=dev)}
    sets.append(d.bind(F, x, y, 1.0, out=out)[0])
torch.cuda.synchronize()
for i in range(AT + 12):
    sets[i % nsets]()
d.synchronize()
t = np.fromfile("/tmp/trace_pair.bin", dtype=np.int64).astype(np.float64)
n = len(t) // 2
names = ["entry", "loads_done", "phaseA_done", "B_start(after wait)", "B_done", "published", "end", "dep_resolved"]
order = [0, 7, 1, 2, 3, 4, 5, 6]
secs = []
for sec in range(2):
    q = t[sec * n:(sec + 1) * n].reshape(-1, 8)
    edge, tiles = q[:WPC], q[WPC:]
    tiles = tiles[(tiles[:, 0] > 0) & (tiles[:, 2] > 0)]
    secs.append((edge, tiles))
t0 = secs[0][1][:, 7].min()     # first warp of launch K-1 past its grid dependency
for sec, (edge, tiles) in enumerate(secs):
    print(f"launch {'K-1' if sec == 0 else 'K'}: {len(tiles)} tile warps; times in us after the first warp of launch K-1 passed griddepcontrol.wait")
    for k in order:
        v = (tiles[:, k] - t0) / 1e3
        print(f"  {names[k]:20s} min {v.min():7.2f} p50 {np.median(v):7.2f} p90 {np.percentile(v, 90):7.2f} max {v.max():7.2f}")
    e = edge[edge[:, 0] > 0]
    print(f"  edge CTA: entry {(e[:, 0].min() - t0) / 1e3:.2f} work done {(e[:, 4].max() - t0) / 1e3:.2f} end {(e[:, 6].max() - t0) / 1e3:.2f}")
a, b = secs
endA = max(a[1][:, 6].max(), a[0][:, 6].max())
print(f"launch K-1: last warp ends {(endA - t0) / 1e3:.2f}; launch K: dependency resolved p50 {(np.median(b[1][:, 7]) - t0) / 1e3:.2f} "
      f"(gap {(np.median(b[1][:, 7]) - endA) / 1e3:.2f} us), last warp ends {(max(b[1][:, 6].max(), b[0][:, 6].max()) - t0) / 1e3:.2f}")
print(f"period (dep_resolved K - dep_resolved K-1): {(np.median(b[1][:, 7]) - np.median(a[1][:, 7])) / 1e3:.2f} us")
for a_, b_ in ((7, 1), (1, 2), (2, 3), (3, 4), (4, 6)):
    v = (b[1][:, b_] - b[1][:, a_]) / 1e3
    print(f"  K: {names[a_]:>20s} -> {names[b_]:20s} p50 {np.median(v):6.2f} p90 {np.percentile(v, 90):6.2f} max {v.max():6.2f}")

# per-SM view of launch K (slot 5 holds %smid)
tk = b[1]
sm = tk[:, 5].astype(int)
dep = np.median(tk[:, 7])
life = (tk[:, 6] - dep) / 1e3
cnt = np.bincount(sm, minlength=148)
print("warps per SM: min", cnt[cnt > 0].min(), "max", cnt.max(), "SMs used", int((cnt > 0).sum()))
for n in sorted(set(cnt[cnt > 0])):
    sel = np.isin(sm, np.where(cnt == n)[0])
    print(f"  SMs with {n:2d} tile warps: {int((cnt == n).sum()):3d} SMs; warp end after dep_resolved: p50 {np.median(life[sel]):5.2f} p90 {np.percentile(life[sel], 90):5.2f} max {life[sel].max():5.2f}")
sm_last = np.array([life[sm == q].max() if (sm == q).any() else 0 for q in range(148)])
print("last warp per SM (us after dep_resolved): min", sm_last[sm_last > 0].min().round(2), "p50", np.median(sm_last[sm_last > 0]).round(2), "max", sm_last.max().round(2))
slow = np.argsort(-sm_last)[:8]
print("slowest SMs:", [(int(q), int(cnt[q]), float(sm_last[q].round(2))) for q in slow])

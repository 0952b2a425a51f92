"""Debug: per-warp timeline of one launch (needs a -DCTD_TRACE build).
usage: trace_run.py FLAGS WPC [problem scheme N]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
F = int(sys.argv[1]); WPC = int(sys.argv[2])
import ctdirect_b200 as ctd
PROB = sys.argv[3] if len(sys.argv) > 3 else "goddard"
SCH = sys.argv[4] if len(sys.argv) > 4 else "midpoint"
NN = int(sys.argv[5]) if len(sys.argv) > 5 else 100000
d = ctd.DOCP(PROB, NN, SCH, stream="own")
dev = torch.device("cuda", 0)
rng = np.random.default_rng(0)
x = torch.from_numpy(d.initial_guess() * (1 + 1e-3 * rng.uniform(-1, 1, d.nvar))).to(dev)
y = torch.from_numpy(rng.standard_normal(d.ncon)).to(dev)
flush = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64, device=dev)
for it in range(4):
    flush.fill_(float(it)); torch.cuda.synchronize()
    if it == 3: os.environ["CTD_TRACE_FILE"] = "/tmp/trace.bin"
    d.evaluate(F, x, y, 1.0); d.synchronize()
t = np.fromfile("/tmp/trace.bin", dtype=np.int64).reshape(-1, 8).astype(np.float64)
edge = t[:WPC]; tiles = t[WPC:]
tiles = tiles[tiles[:, 0] > 0]
t0 = min(tiles[:, 0].min(), edge[edge[:, 0] > 0][:, 0].min())
print("warps traced", len(tiles), " kernel span (first start -> last end) us:", (tiles[:, 6].max() - t0) / 1e3)
names = ["start", "loads_done", "phaseA_done", "B_start(after wait)", "B_done", "published", "end"]
for k in range(7):
    v = (tiles[:, k] - t0) / 1e3
    print(f"  {names[k]:20s} min {v.min():7.2f} p50 {np.median(v):7.2f} p90 {np.percentile(v,90):7.2f} max {v.max():7.2f}")
dur = (tiles[:, 6] - tiles[:, 0]) / 1e3
print("  warp lifetime us: p50", np.median(dur), "p90", np.percentile(dur, 90), "max", dur.max())
for a_, b_ in ((0, 1), (1, 2), (2, 3), (3, 4), (4, 5), (5, 6)):
    v = (tiles[:, b_] - tiles[:, a_]) / 1e3
    print(f"  {names[a_]:>18s} -> {names[b_]:18s} p50 {np.median(v):6.2f} p90 {np.percentile(v,90):6.2f} max {v.max():6.2f}")
e = edge[edge[:, 0] > 0]
print("  edge CTA: start", (e[:, 0].min() - t0) / 1e3, "edge work done", (e[:, 4].max() - t0) / 1e3, "end", (e[:, 6].max() - t0) / 1e3)
late = tiles[:, 0] - t0 > 3000
print("  warps starting later than 3us after first:", int(late.sum()))

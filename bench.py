#!/usr/bin/env python
"""Benchmark of the DOCP evaluation path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

One "step" = one fused cons + Jacobian + Hessian evaluation (ctd_eval_dev, flags CONS|JAC|HESS) of the
Goddard / midpoint / grid_size=1e5 DOCP on synthetic inputs resident in HBM.  With N > 1 ranks (torchrun)
every rank evaluates its own instance (independent scenarios, no data-path collective): weak scaling,
`value` = evaluations of all ranks / max-over-ranks device time.  The same JSON line carries
  e2e           the same evaluation through the host-pointer C ABI call (ctd_eval) with pinned host buffers,
                H2D of x,y and D2H of c, J values, H values inside the timed region;
  roofline      algorithmic bytes / measured launch time against the measured HBM copy bandwidth;
  cpu_baseline  the CPU oracle port (oracle/, test infrastructure) timed on the host cores;
  batched       the 4096-instance Goddard MPC-scenario batch (grid_size=1000) sharded over the ranks with
                one NCCL all-gather of (objective, violation) per evaluation.
`--impl reference` times the CPU port of the reference's algorithm instead (the reference itself is Julia
and cannot run in this image).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

PROBLEM, SCHEME, GRID = "goddard", "midpoint", 100_000
B_TOTAL, B_GRID = 4096, 1000
SEED = 20261018
METRIC = "cons+Jacobian+Hessian evals/s @ grid 1e5"


def synthetic_point(x0, ncon, seed, batch=1):
    """reference initial guess perturbed by 1 + 1e-3*U(-1,1); y ~ N(0,1)."""
    rng = np.random.default_rng(seed)
    x = x0[None, :] * (1.0 + 1e-3 * rng.uniform(-1.0, 1.0, size=(batch, x0.size)))
    y = rng.standard_normal((batch, ncon))
    return np.ascontiguousarray(x), np.ascontiguousarray(y)


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """Samples SM clock and throttle reasons with NVML while a timed region runs."""
    REASONS = {
        0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown",
        0x10: "sync_boost", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown", 0x80: "hw_power_brake_slowdown",
        0x100: "display_clock_setting",
    }

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nv = None

    def _loop(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                self.samples.append(int(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                mask = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                for bit, name in self.REASONS.items():
                    if mask & bit and name != "gpu_idle":
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(0.005)

    def start(self):
        if self.nv:
            self._stop.clear()
            self._thr = threading.Thread(target=self._loop, daemon=True)
            self._thr.start()

    def stop(self):
        if self._thr:
            self._stop.set()
            self._thr.join()
            self._thr = None

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": 0}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ---------------------------------------------------------------------------------------------------
def cpu_fused_eval_time(orc_mod, n_threads, budget_s, min_reps=2):
    """Time one fused cons + Jacobian + Hessian evaluation of the CPU oracle at the bench workload."""
    o = orc_mod.OracleModel(PROBLEM, SCHEME, GRID)
    x, y = synthetic_point(o.x0(), o.ncon, SEED)
    x, y = x[0], y[0]
    used = orc_mod.set_threads(n_threads)
    o.cons(x)
    o.derivs(x, y, 1.0, want=("jac", "hess"))
    reps, t0 = 0, time.perf_counter()
    while True:
        o.cons(x)
        o.derivs(x, y, 1.0, want=("jac", "hess"))
        reps += 1
        el = time.perf_counter() - t0
        if reps >= min_reps and el >= budget_s:
            break
        if reps >= 200:
            break
    return el / reps, reps, used


def workload_text(nvar, ncon, nnzj, nnzh):
    """`config.workload`, identical in both arms."""
    return (f"{PROBLEM}/{SCHEME}/grid_size={GRID}: fused cons+jac+hess, :manual pattern, "
            f"nvar={nvar} ncon={ncon} nnzj={nnzj} nnzh={nnzh}; one instance per GPU")


def run_reference(args, rank):
    """--impl reference: the CPU port of the reference's algorithm on the host cores (rank 0 only)."""
    if rank != 0:
        return
    from oracle import oracle as orc
    orc.build()
    ncores = os.cpu_count() or 1
    o = orc.OracleModel(PROBLEM, SCHEME, GRID)
    x, y = synthetic_point(o.x0(), o.ncon, SEED)
    x, y = x[0], y[0]
    used = orc.set_threads(ncores)
    for _ in range(max(args.warmup, 1)):
        o.cons(x)
        o.derivs(x, y, 1.0, want=("jac", "hess"))
    steps = max(1, min(args.steps, 50))
    t0 = time.perf_counter()
    for _ in range(steps):
        o.cons(x)
        o.derivs(x, y, 1.0, want=("jac", "hess"))
    el = time.perf_counter() - t0
    val = steps / el
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": "evals/s", "n_gpus": args.gpus,
        "steps": steps, "warmup": max(args.warmup, 1), "ms_per_step": 1e3 * el / steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_text(o.nvar, o.ncon, o.nnzj, o.nnzh),
                   "note": "CPU port of the reference algorithm (oracle/): the reference is Julia + ADNLPModels and "
                           "cannot run in this image; derivatives by element-wise second-order forward AD, OpenMP"},
        "cpu_baseline": {"value": val, "unit": "evals/s", "cores": used, "kind": "port",
                         "sample": f"{steps} fused evaluations of the full workload"},
        "e2e": {"value": val, "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    _emit(line)


# ---------------------------------------------------------------------------------------------------
_REAL_STDOUT = None


def _quiet_stdout():
    """Route file descriptor 1 to stderr while the benchmark runs (NCCL / torch print banners on stdout);
    the single JSON line is written to the real stdout by _emit()."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def _emit(line: dict):
    data = (json.dumps(line) + "\n").encode()
    sys.stdout.flush()
    if _REAL_STDOUT is not None:
        os.write(_REAL_STDOUT, data)
    else:
        sys.stdout.write(data.decode())
        sys.stdout.flush()


def main():
    _quiet_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--streams", type=int, default=3, help="independent evaluation streams in flight")
    ap.add_argument("--no-batched", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank)
        return 0
    if args.warmup < 3:
        args.warmup = 3

    import torch
    import torch.distributed as dist
    import ctdirect_b200 as ctd

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback); use --impl reference for the CPU port")
    numa = None
    if world > 1 and not os.environ.get("CTD_BENCH_NO_NUMA"):   # one process per GPU: sit on the GPU's NUMA node before any pinned buffer is allocated
        from ctdirect_b200.batch import bind_to_gpu_numa_node
        numa = bind_to_gpu_numa_node(local_rank)
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"   # keep stdout to the single JSON line
        os.environ.setdefault("TORCH_NCCL_HIGH_PRIORITY", "1")   # the gather's CTAs go ahead of queued evaluation CTAs
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        return ms

    FLAGS = ctd.CONS | ctd.JAC | ctd.HESS
    # Evaluations are independent of each other (MPC scenarios, trial points): they are issued round-robin
    # on NSTREAMS handles, each with its own CUDA stream, so that the launch latency and the tail of one
    # evaluation overlap the body of the next.  The single-stream figure is reported beside it.
    NSTREAMS = max(1, args.streams)
    docps = [ctd.DOCP(PROBLEM, GRID, SCHEME, device=local_rank) for _ in range(NSTREAMS)]
    docp = docps[0]
    nlp = docp.nlp_model()
    nvar, ncon, nnzj, nnzh = docp.nvar, docp.ncon, docp.nnzj, docp.nnzh
    alg_bytes = 8 * (nvar + 2 * ncon + nnzj + nnzh)
    streams = [torch.cuda.ExternalStream(d.cuda_stream, device=dev) for d in docps]
    stream = streams[0]
    # rotating buffer sets so that consecutive evaluations do not find their lines in the 126 MB L2
    nsets = max(3, int(np.ceil(1.5 * 126e6 / alg_bytes)))
    nsets = ((nsets + NSTREAMS - 1) // NSTREAMS) * NSTREAMS
    xh, yh = synthetic_point(nlp.meta.x0, ncon, SEED + rank)
    sets = []
    for s in range(nsets):
        xs, ys = synthetic_point(nlp.meta.x0, ncon, SEED + 1000 * s + rank)
        sets.append({
            "x": torch.from_numpy(xs[0]).to(dev), "y": torch.from_numpy(ys[0]).to(dev),
            "out": {"cons": torch.empty(ncon, dtype=torch.float64, device=dev),
                    "jac": torch.empty(nnzj, dtype=torch.float64, device=dev),
                    "hess": torch.empty(nnzh, dtype=torch.float64, device=dev)},
        })
    torch.cuda.synchronize()
    # buffer set k is always evaluated by handle k % NSTREAMS
    calls = [docps[k % NSTREAMS].bind(FLAGS, s["x"], s["y"], 1.0, out=s["out"])[0] for k, s in enumerate(sets)]
    calls1 = [docp.bind(FLAGS, s["x"], s["y"], 1.0, out=s["out"])[0] for s in sets]

    REPEATS = 5

    def timed(call_list, used_streams, steps, warmup):
        """REPEATS timed regions of exactly `steps` evaluations each (barrier + synchronize on both sides of
        every region); returns the per-region times (ms, this rank) and the launches of one region.  A region
        of 100 evaluations lasts under a millisecond, so a single region is at the mercy of one host hiccup
        on any rank: the reported figure is the median region of the max over ranks."""
        for i in range(warmup):
            call_list[i % nsets]()
        regions, launches = [], 0
        for _ in range(REPEATS):
            barrier()
            l0 = sum(d.launch_count for d in docps)
            ev0 = [torch.cuda.Event(enable_timing=True) for _ in used_streams]
            ev1 = [torch.cuda.Event(enable_timing=True) for _ in used_streams]
            for e, st in zip(ev0, used_streams):
                e.record(st)
            for i in range(steps):
                call_list[i % nsets]()
            for e, st in zip(ev1, used_streams):
                e.record(st)
            barrier()
            # every stream starts at the barrier; the region ends when the last stream finishes
            regions.append(max(a.elapsed_time(b) for a, b in zip(ev0, ev1)))
            launches = sum(d.launch_count for d in docps) - l0
        return regions, launches

    def regions_over_ranks(regions):
        """max over ranks of every region -> (median region, all regions)"""
        t = torch.tensor(regions, dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        r = sorted(t.tolist())
        return r[len(r) // 2], r

    sampler = ClockSampler(local_rank)
    # (1) single stream: the kernel timed back to back on its own -> roofline of the kernel
    reg1_local, _ = timed(calls1, [stream], args.steps, args.warmup)
    # (2) NSTREAMS handles in flight -> whole-job throughput (`value`)
    sampler.start()
    reg_local, launches = timed(calls, streams, args.steps, args.warmup)
    ms, regions = regions_over_ranks(reg_local)
    ms1, regions1 = regions_over_ranks(reg1_local)
    ms_local = sorted(reg_local)[len(reg_local) // 2]
    ms1_local = sorted(reg1_local)[len(reg1_local) // 2]
    value = world * args.steps / (ms * 1e-3)
    peak, peak_src = measured_peak()
    launch_s = ms1_local * 1e-3 / args.steps
    achieved = alg_bytes / launch_s / 1e9
    achieved_ov = alg_bytes / (ms_local * 1e-3 / args.steps) / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        try:
            traffic = json.load(open(tp)).get("dram_bytes_per_launch")
        except Exception:
            traffic = None
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "kernel": "ctd_eval_kernel_light<E_goddard__midpoint, CONS|JAC|HESS>",
                "algorithmic_bytes_per_launch": alg_bytes, "launch_us": launch_s * 1e6, "peak_source": peak_src,
                "how": "one stream, launches back to back, CUDA events over the timed region / steps",
                "overlapped": {"streams": NSTREAMS, "achieved": achieved_ov, "frac": achieved_ov / peak,
                               "us_per_eval": ms_local * 1e3 / args.steps,
                               "how": "same kernel, evaluations issued round-robin on independent streams"}}
    single_stream = {"value": world * args.steps / (ms1 * 1e-3), "unit": "evals/s", "ms_per_step": ms1 / args.steps}

    # ---- e2e: host buffers through the host-pointer C ABI (H2D + kernel + D2H every step).
    # The evaluations are independent, so NSTREAMS of them are kept in flight (ctd_eval_async on one handle
    # each, ctd_sync before a handle's buffers are reused): the H2D, kernel and D2H stages of consecutive
    # evaluations overlap on the full-duplex link.  The strictly synchronous figure is reported beside it.
    e2e = None
    if not args.no_e2e:
        pin = lambda n: torch.empty(n, dtype=torch.float64).pin_memory()
        hsets = []
        for k in range(NSTREAMS):
            hx, hy = pin(nvar), pin(ncon)
            hx.numpy()[:] = xh[0]
            hy.numpy()[:] = yh[0]
            hout = {"cons": pin(ncon).numpy(), "jac": pin(nnzj).numpy(), "hess": pin(nnzh).numpy()}
            hsets.append((hx, hy, hout))
        sync_call = docp.bind(FLAGS, hsets[0][0].numpy(), hsets[0][1].numpy(), 1.0, out=hsets[0][2])[0]
        async_calls = [docps[k].bind(FLAGS, hsets[k][0].numpy(), hsets[k][1].numpy(), 1.0, out=hsets[k][2], wait=False)[0]
                       for k in range(NSTREAMS)]
        e2e_steps = max(3, min(args.steps, 300))
        for _ in range(3):
            sync_call()
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            sync_call()   # synchronous: returns with results on the host
        el_sync = time.perf_counter() - t0
        def pipelined_region():
            barrier()
            t0 = time.perf_counter()
            for i in range(e2e_steps):
                k = i % NSTREAMS
                if i >= NSTREAMS:
                    docps[k].synchronize()          # results of evaluation i - NSTREAMS are on the host
                async_calls[k]()
            for d_ in docps:
                d_.synchronize()
            return max_over_ranks((time.perf_counter() - t0) * 1e3) * 1e-3

        el_pipe = sorted(pipelined_region() for _ in range(3))[1]   # median of 3 regions (wall clock is noisy)
        el_sync = max_over_ranks(el_sync * 1e3) * 1e-3
        legs = {"pipelined": {"value": world * e2e_steps / el_pipe, "ms_per_step": 1e3 * el_pipe / e2e_steps,
                              "api": f"ctd_eval_async (host pointers, pinned buffers), {NSTREAMS} evaluations in flight"},
                "synchronous": {"value": world * e2e_steps / el_sync, "ms_per_step": 1e3 * el_sync / e2e_steps,
                                "api": "ctd_eval: one blocking call per evaluation"}}
        best = max(legs, key=lambda k_: legs[k_]["value"])
        # both legs move the same bytes over PCIe (the D2H of 42 MB bounds them); the headline is the faster one
        e2e = {"value": legs[best]["value"], "unit": "evals/s", "h2d_bytes_per_step": 8 * (nvar + ncon),
               "d2h_bytes_per_step": 8 * (ncon + nnzj + nnzh), "steps": e2e_steps,
               "ms_per_step": legs[best]["ms_per_step"], "api": legs[best]["api"] + ", wall clock", "mode": best,
               "pipelined": legs["pipelined"], "synchronous": legs["synchronous"]}
    sampler.stop()

    # ---- batched MPC scenarios: 4096 instances of grid 1000 sharded over the ranks
    batched = None
    if not args.no_batched:
        bl = B_TOTAL // world
        rng = np.random.default_rng(SEED + 17)
        tmax = rng.uniform(1.0, 3.5, size=B_TOTAL)          # continuation range test_initial_guess.jl:242
        par = np.tile(np.array([310.0, 500.0, 2.0, 3.5]), (B_TOTAL, 1))
        par[:, 3] = tmax
        par = par[rank * bl:(rank + 1) * bl]
        bd = ctd.DOCP(PROBLEM, B_GRID, SCHEME, params=par, batch=bl, device=local_rank)
        bx, by = synthetic_point(bd.initial_guess(), bd.ncon, SEED + 99 + rank, batch=bl)
        tx, ty = torch.from_numpy(bx).to(dev), torch.from_numpy(by).to(dev)
        bout = {"cons": torch.empty((bl, bd.ncon), dtype=torch.float64, device=dev),
                "jac": torch.empty((bl, bd.nnzj), dtype=torch.float64, device=dev),
                "hess": torch.empty((bl, bd.nnzh), dtype=torch.float64, device=dev),
                "grad": torch.empty((bl, bd.nvar), dtype=torch.float64, device=dev),
                "obj": torch.empty(bl, dtype=torch.float64, device=dev)}
        norms = torch.empty((2, bl), dtype=torch.float64, device=dev)
        gathered = torch.empty((world * 2, bl), dtype=torch.float64, device=dev)
        ALL = ctd.CONS | ctd.JAC | ctd.HESS | ctd.GRAD | ctd.OBJ
        bstream = torch.cuda.ExternalStream(bd.cuda_stream, device=dev)
        bsteps = max(3, min(args.steps, 50))

        bout["obj"] = norms[0]     # objective and violation land directly in the buffer that is gathered

        beval = bd.bind_with_norms(ALL, tx, ty, 1.0, out=bout, viol=norms[1])[0]

        def bstep():
            beval()
            if world > 1:
                dist.all_gather_into_tensor(gathered, norms)

        def timed_batched(step, drain=None):
            regs = []
            with torch.cuda.stream(bstream):
                for _ in range(3):
                    step()
                for _ in range(3):          # 3 timed regions of bsteps evaluations; the median is reported
                    if drain:
                        drain()
                    barrier()
                    b0, b1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    b0.record(bstream)
                    for _ in range(bsteps):
                        step()
                    if drain:
                        drain()             # the timed region ends when the last gather has landed
                    b1.record(bstream)
                    barrier()
                    regs.append(b0.elapsed_time(b1))
            return regions_over_ranks(regs)[0]

        bms = timed_batched(bstep)
        # the same with the all-gather of evaluation k overlapping the kernel of evaluation k+1
        # (ctdirect_b200.batch.NormsPipeline: collective on its own stream, two result slots)
        pms = None
        if world > 1:
            from ctdirect_b200.batch import NormsPipeline
            pipe = NormsPipeline(bl, dev, depth=2, eval_stream=bstream)

            pevals = []
            for slot in range(2):
                po = dict(bout)
                po["obj"] = pipe.local[slot][0]
                pevals.append(bd.bind_with_norms(ALL, tx, ty, 1.0, out=po, viol=pipe.local[slot][1])[0])

            def pstep():
                return pipe.submit(lambda slot, o, v: pevals[slot]())

            pms = timed_batched(pstep, pipe.drain)
        # the same two schedules through the library's own communicator (ctd_comm_init / ctd_gather_norms_dev:
        # two C calls per step instead of a torch.distributed call)
        nms = nps = None
        if world > 1:
            from ctdirect_b200.batch import NativeNormsPipeline, broadcast_unique_id
            npipe = NativeNormsPipeline(bd, rank, world, broadcast_unique_id(), depth=2, n_instances=bl * world)
            nout = {k_: v_ for k_, v_ in bout.items() if k_ != "obj"}
            nms = timed_batched(lambda: npipe.submit(ALL, tx, ty, nout, serial=True), npipe.drain)
            nps = timed_batched(lambda: npipe.submit(ALL, tx, ty, nout), npipe.drain)
        bbytes = 8 * bl * (2 * bd.nvar + 2 * bd.ncon + bd.nnzj + bd.nnzh)
        legs_b = {"kernel then all-gather on one stream (torch.distributed)": bms,
                  "all-gather of evaluation k overlaps the kernel of evaluation k+1 (torch.distributed, separate stream, 2 slots)": pms,
                  "kernel then all-gather, both issued by the library (ctd_gather_norms_dev)": nms,
                  "all-gather of evaluation k overlaps the kernel of evaluation k+1, both issued by the library "
                  "(ctd_gather_norms_dev on the handle's communication stream, 2 slots)": nps}
        best_how = min((k_ for k_ in legs_b if legs_b[k_] is not None), key=lambda k_: legs_b[k_])
        best = legs_b[best_how]
        bach = bbytes / (best * 1e-3 / bsteps) / 1e9
        batched = {"workload": f"{B_TOTAL} Goddard/midpoint instances, grid_size={B_GRID}, Tmax in [1,3.5], "
                               f"{bl} per GPU; obj+grad+cons+jac+hess + violation norm in ONE launch + NCCL all-gather",
                   "batched_evals_per_s": bsteps / (best * 1e-3), "instance_evals_per_s": B_TOTAL * bsteps / (best * 1e-3),
                   "ms_per_batched_eval": best / bsteps, "steps": bsteps, "scaling": "strong",
                   "hbm_gbs_per_gpu": bach, "hbm_frac_per_gpu": bach / peak,
                   "how": best_how,
                   "serialised": {"batched_evals_per_s": bsteps / (bms * 1e-3), "ms_per_batched_eval": bms / bsteps,
                                  "how": "kernel then all-gather on one stream (torch.distributed)"},
                   "native_serialised": (None if nms is None else
                                         {"batched_evals_per_s": bsteps / (nms * 1e-3), "ms_per_batched_eval": nms / bsteps}),
                   "native_pipelined": (None if nps is None else
                                        {"batched_evals_per_s": bsteps / (nps * 1e-3), "ms_per_batched_eval": nps / bsteps}),
                   "pipelined": (None if pms is None else
                                 {"batched_evals_per_s": bsteps / (pms * 1e-3), "ms_per_batched_eval": pms / bsteps})}

    # ---- CPU baseline on the host cores (rank 0, bounded sample)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        try:
            from oracle import oracle as orc
            orc.build()
            ncores = os.cpu_count() or 1
            t1, r1, _ = cpu_fused_eval_time(orc, 1, 4.0)
            tn, rn, used = cpu_fused_eval_time(orc, ncores, 8.0)
            cpu = {"value": 1.0 / tn, "unit": "evals/s", "cores": used, "kind": "port",
                   "sample": f"{rn} fused cons+jac+hess evaluations of the full workload on {used} threads "
                             f"(OpenMP over steps); single thread: {1.0 / t1:.3f} evals/s over {r1} evaluations",
                   "single_thread_value": 1.0 / t1}
        except Exception as ex:  # the oracle is only a reported baseline
            cpu = {"value": None, "unit": "evals/s", "cores": 0, "kind": "port", "sample": f"failed: {ex}"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "evals/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_text(nvar, ncon, nnzj, nnzh),
                       "launches": "one kernel launch per evaluation",
                       "l2": f"rotating {nsets} input/output buffer sets ({nsets * alg_bytes / 1e6:.0f} MB > 126 MB L2)",
                       "streams": f"{NSTREAMS} independent evaluations in flight per GPU (one handle + CUDA stream each); "
                                  "single_stream / roofline give the one-stream figures",
                       "timing": f"CUDA events on the handles' streams, max over streams and ranks; median of {REPEATS} timed regions of exactly --steps evaluations each (all regions in region_ms)"},
            "single_stream": single_stream, "region_ms": regions, "numa": numa,
            "clocks": sampler.summary(), "e2e": e2e, "gpu_launches": int(launches),
            "roofline": roofline, "cpu_baseline": cpu, "batched": batched,
        }
        _emit(line)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())

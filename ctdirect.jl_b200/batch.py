"""Batched MPC scenarios sharded over the GPUs of one node (BASELINE config 5).

Independent instances of the same DOCP (different parameters / initial states) are split into
contiguous blocks, one block per rank (one process per GPU).  The data path needs no collective; the only
exchange is one all-gather of the per-instance (objective, constraint violation) pairs, 2*B doubles per
evaluation, over NCCL (GPU tensors) or gloo (CPU tensors, used by the CPU tests).
The reference has no batching at all (every scenario would be its own DOCP)."""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np


def shard_range(n_instances: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Contiguous block [lo, hi) of instances owned by `rank`; blocks differ by at most one instance."""
    if world_size < 1 or not (0 <= rank < world_size):
        raise ValueError("bad world_size / rank")
    base, rem = divmod(n_instances, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_norms(obj, viol, n_instances: int, group=None):
    """All-gather the local (objective, violation) vectors into two global vectors of n_instances entries,
    in instance order, on every rank.  `obj` and `viol` are torch tensors (CUDA -> NCCL, CPU -> gloo)."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return obj.clone(), viol.clone()
    world = dist.get_world_size(group)
    sizes = [shard_range(n_instances, world, r) for r in range(world)]
    nmax = max(hi - lo for lo, hi in sizes)
    local = torch.zeros((2, nmax), dtype=torch.float64, device=obj.device)
    local[0, :obj.numel()] = obj
    local[1, :viol.numel()] = viol
    out = torch.empty((world * 2, nmax), dtype=torch.float64, device=obj.device)
    dist.all_gather_into_tensor(out, local, group=group)   # concatenation along dim 0 (NCCL and gloo)
    out = out.view(world, 2, nmax)
    g_obj = torch.cat([out[r, 0, :hi - lo] for r, (lo, hi) in enumerate(sizes)])
    g_viol = torch.cat([out[r, 1, :hi - lo] for r, (lo, hi) in enumerate(sizes)])
    return g_obj, g_viol


class ShardedScenarios:
    """The local shard of a batch of scenarios: a batched DOCP handle for instances [lo, hi)."""

    def __init__(self, ocp, n_instances: int, params: np.ndarray, grid_size: int = 1000, scheme: str = "midpoint",
                 rank: int = 0, world_size: int = 1, device: Optional[int] = None, **kwargs):
        from .docp import DOCP
        params = np.ascontiguousarray(params, dtype=np.float64)
        if params.ndim != 2 or params.shape[0] != n_instances:
            raise ValueError("params must have one row per instance")
        self.n_instances, self.rank, self.world_size = n_instances, rank, world_size
        self.lo, self.hi = shard_range(n_instances, world_size, rank)
        self.docp = DOCP(ocp, grid_size=grid_size, scheme=scheme, params=params[self.lo:self.hi],
                         batch=self.hi - self.lo, device=rank if device is None else device, **kwargs)

    def norms(self, x, group=None):
        """Local evaluation of (objective, violation) followed by the all-gather: global vectors."""
        obj, viol = self.docp.norms(x)
        return gather_norms(obj, viol, self.n_instances, group)


class NormsPipeline:
    """Overlaps the all-gather of batched evaluation k with the kernel of evaluation k+1.

    At 8 GPUs the local evaluation of 512 instances takes about 48 us and the NCCL all-gather of the
    2*B doubles about 25 us of pure latency; issued back to back on one stream they add up.  Here the
    collective runs on its own stream behind an event, `depth` result slots deep: `submit` enqueues
    evaluation k (its objective / violation vectors land directly in the slot that is gathered) and the
    gather of that slot; `result(slot)` makes the caller's stream wait for that gather and returns the
    (world, 2, n_local) tensor.  With CPU tensors (gloo, tests) there are no streams and every step is
    serial; the results are identical.

    `evaluate(slot, obj, viol)` is the user's evaluation: it must write the local objective and violation
    vectors into the two given rows (e.g. `DOCP.evaluate_with_norms(..., out={"obj": obj}, viol=viol)`)."""

    def __init__(self, n_local: int, device, group=None, depth: int = 2, eval_stream=None):
        import torch
        import torch.distributed as dist
        self.group = group
        self.world = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
        self.depth = depth
        self.cuda = torch.device(device).type == "cuda"
        self.local = [torch.zeros((2, n_local), dtype=torch.float64, device=device) for _ in range(depth)]
        self.gathered = [torch.zeros((self.world * 2, n_local), dtype=torch.float64, device=device) for _ in range(depth)]
        self.k = 0
        if self.cuda:
            self.eval_stream = eval_stream if eval_stream is not None else torch.cuda.current_stream(device)
            self.comm_stream = torch.cuda.Stream(device=device)
            self.ev_done = [torch.cuda.Event() for _ in range(depth)]
            self.ev_gathered = [torch.cuda.Event() for _ in range(depth)]
            self.used = [False] * depth

    def submit(self, evaluate) -> int:
        """Enqueue one evaluation and the all-gather of its norms; returns the slot holding the result."""
        import torch
        import torch.distributed as dist
        s = self.k % self.depth
        self.k += 1
        if not self.cuda:
            evaluate(s, self.local[s][0], self.local[s][1])
            if self.world > 1:
                dist.all_gather_into_tensor(self.gathered[s], self.local[s], group=self.group)
            else:
                self.gathered[s].copy_(self.local[s])
            return s
        if self.used[s]:   # the slot's previous gather must have read the local vectors
            self.eval_stream.wait_event(self.ev_gathered[s])
        with torch.cuda.stream(self.eval_stream):
            evaluate(s, self.local[s][0], self.local[s][1])
        self.ev_done[s].record(self.eval_stream)
        self.comm_stream.wait_event(self.ev_done[s])
        with torch.cuda.stream(self.comm_stream):
            if self.world > 1:
                dist.all_gather_into_tensor(self.gathered[s], self.local[s], group=self.group)
            else:
                self.gathered[s].copy_(self.local[s], non_blocking=True)
        self.ev_gathered[s].record(self.comm_stream)
        self.used[s] = True
        return s

    def result(self, slot: int, stream=None):
        """The gathered norms of `slot` as a (world, 2, n_local) tensor; `stream` (default: the evaluation
        stream) is made to wait for the collective."""
        if self.cuda and self.used[slot]:
            (stream if stream is not None else self.eval_stream).wait_event(self.ev_gathered[slot])
        return self.gathered[slot].view(self.world, 2, -1)

    def drain(self):
        """Make the evaluation stream wait for every outstanding gather."""
        if self.cuda:
            for s in range(self.depth):
                if self.used[s]:
                    self.eval_stream.wait_event(self.ev_gathered[s])


def broadcast_unique_id(group=None) -> bytes:
    """Create the NCCL id of the library's communicator on rank 0 and hand it to every rank through the
    existing torch.distributed group (any backend)."""
    import torch.distributed as dist
    from .docp import DOCP
    box = [DOCP.nccl_unique_id() if dist.get_rank(group) == 0 else None]
    dist.broadcast_object_list(box, src=0, group=group)
    return box[0]


class NativeNormsPipeline:
    """`NormsPipeline` on the library's own communicator: the evaluation and the all-gather are two C calls
    per step (ctd_eval_norms_dev, ctd_gather_norms_dev), the collective runs on a high-priority stream owned
    by the handle, and slot events replace the Python-side stream bookkeeping.  Shards must have the same
    size on every rank.

        pipe = NativeNormsPipeline(docp, rank, world, unique_id)
        slot = pipe.submit(what, x, y, out)        # evaluation k and its gather (overlaps evaluation k+1)
        norms = pipe.result(slot)                  # (world, 2, batch) tensor, valid on the handle's stream
    """

    def __init__(self, docp, rank: int, world: int, unique_id: bytes, depth: int = 2, n_instances: Optional[int] = None):
        import torch
        if not (1 <= depth <= 4):
            raise ValueError("depth must be 1..4 (CTD_COMM_SLOTS)")
        # ncclAllGather moves the same count from every rank: unequal shards would corrupt the result or hang
        if n_instances is not None and docp.batch * world != n_instances:
            raise ValueError(f"shards must have the same size on every rank: {n_instances} instances over {world} ranks "
                             f"with {docp.batch} here (pad the batch to a multiple of the number of ranks)")
        self.docp, self.world, self.depth, self.k = docp, world, depth, 0
        docp.comm_init(unique_id, rank, world)
        dev = torch.device("cuda", docp.device)
        self.local = [torch.zeros((2, docp.batch), dtype=torch.float64, device=dev) for _ in range(depth)]
        self.gathered = [torch.zeros((world, 2, docp.batch), dtype=torch.float64, device=dev) for _ in range(depth)]
        self._bound = {}

    def submit(self, what: int, x, y=None, out=None, obj_weight: float = 1.0, serial: bool = False) -> int:
        s = self.k % self.depth
        self.k += 1
        key = (s, what, x.data_ptr(), 0 if y is None else y.data_ptr(), obj_weight,
               tuple(sorted((k, v.data_ptr()) for k, v in (out or {}).items() if k != "obj")))
        call = self._bound.get(key)
        if call is None:   # buffers are validated once per distinct set
            o = dict(out or {})
            o["obj"] = self.local[s][0]
            call = self.docp.bind_with_norms(what, x, y, obj_weight, out=o, viol=self.local[s][1])[0]
            self._bound[key] = call
        self.docp.comm_wait(s)          # the gather that last used this slot has read local[s]
        call()
        self.docp.gather_norms(self.local[s], self.gathered[s], s)
        if serial:
            self.docp.comm_wait(s)
        return s

    def result(self, slot: int):
        self.docp.comm_wait(slot)
        return self.gathered[slot]

    def drain(self):
        for s in range(self.depth):
            self.docp.comm_wait(s)


class PeerNormsPipeline:
    """The norm gather fused into the evaluation kernel over NVLink peer memory: one C call per step issues the
    evaluation, whose reduction epilogue stores every instance's (objective, violation) into the gather buffer of
    every rank (ctd_eval_gather_dev); `result(slot)` enqueues the wait for all ranks' arrival flags (a one-warp
    kernel) and returns the (world, 2, batch) view.  No collective call, no second stream.

        pipe = PeerNormsPipeline(docp, rank, world, unique_id)
        slot = pipe.submit(what, x, y, out)      # evaluation k; its norms travel while evaluation k+1 runs
        norms = pipe.result(slot)                # valid on the handle's stream
    A slot is reused `depth` steps later; every rank must have waited for it by then (`submit` does, like
    NativeNormsPipeline)."""

    def __init__(self, docp, rank: int, world: int, unique_id: bytes, depth: int = 2, n_instances: Optional[int] = None):
        if not (1 <= depth <= 4):
            raise ValueError("depth must be 1..4 (CTD_COMM_SLOTS)")
        if n_instances is not None and docp.batch * world != n_instances:
            raise ValueError(f"shards must have the same size on every rank: {n_instances} instances over {world} ranks "
                             f"with {docp.batch} here")
        self.docp, self.world, self.depth, self.k = docp, world, depth, 0
        if not getattr(docp, "comm_world", None):
            docp.comm_init(unique_id, rank, world)
        docp.peer_init()
        self._bound = {}
        self._pending = [False] * depth
        self._view = [None] * depth

    def submit(self, what: int, x, y=None, out=None, obj_weight: float = 1.0, serial: bool = False) -> int:
        s = self.k % self.depth
        self.k += 1
        # (no host-side wait for the slot's previous round: the launch itself waits, in its first CTA, until every rank --
        #  this one included -- has copied that round out; whoever reads a slot does so through result() before the
        #  slot comes round again)
        key = (s, what, x.data_ptr(), 0 if y is None else y.data_ptr(), obj_weight,
               tuple(sorted((k, v.data_ptr()) for k, v in (out or {}).items() if k != "obj")))
        call = self._bound.get(key)
        if call is None:
            o = {k: v for k, v in (out or {}).items() if k != "obj"}
            call = self.docp.bind_gather(what, x, y, obj_weight, out=o, slot=s)[0]
            self._bound[key] = call
        call()
        if serial:
            self._view[s] = self.docp.peer_wait(s)                     # on the evaluation stream: what follows sees the norms
        else:
            self._view[s] = self.docp.peer_wait(s, overlapped=True)    # beside the next evaluation
            self._pending[s] = True
        return s

    def result(self, slot: int):
        if self._pending[slot]:
            self.docp.comm_wait(slot)
            self._pending[slot] = False
        return self._view[slot]

    def drain(self):
        for s in range(self.depth):
            if self._pending[s]:
                self.docp.comm_wait(s)
                self._pending[s] = False


class SplitShardPipeline:
    """A rank's shard as SEVERAL handles on their own streams, each with its own fused peer-memory gather.

    A batched evaluation of a small shard (512 instances per GPU when 4096 are spread over 8 GPUs) is a ~45 us kernel
    whose ramp-up and drain cost 10 % of it; two half-shards on two independent streams fill each other's ramp-up and
    drain (8 GPUs: 20.3 k -> 22.5 k batched evaluations per second; four quarter-shards: too many small launches).

        parts = [DOCP(problem, N, scheme, params=par[lo:mid], batch=mid - lo, device=dev, stream="own"),
                 DOCP(problem, N, scheme, params=par[mid:hi], batch=hi - mid, device=dev, stream="own")]
        pipe = SplitShardPipeline(parts, rank, world)
        slot = pipe.submit(what, xs, ys, outs)      # lists, one entry per part
        norms = pipe.result(slot)                   # list of (world, 2, batch_of_part) views, one per part
    Every rank must split its shard the same way (the gathers of part k pair up across ranks)."""

    def __init__(self, parts, rank: int, world: int, depth: int = 4, group=None):
        if not parts:
            raise ValueError("at least one part")
        self.parts = list(parts)
        self.pipes = [PeerNormsPipeline(d, rank, world, broadcast_unique_id(group), depth=depth,
                                        n_instances=d.batch * world) for d in self.parts]

    def submit(self, what: int, xs, ys=None, outs=None, obj_weight: float = 1.0, serial: bool = False) -> int:
        n = len(self.parts)
        ys = ys if ys is not None else [None] * n
        outs = outs if outs is not None else [None] * n
        if not (len(xs) == len(ys) == len(outs) == n):
            raise ValueError(f"one x / y / out per part expected ({n} parts)")
        slots = {p.submit(what, x, y, o, obj_weight, serial) for p, x, y, o in zip(self.pipes, xs, ys, outs)}
        assert len(slots) == 1
        return slots.pop()

    def result(self, slot: int):
        return [p.result(slot) for p in self.pipes]

    def drain(self):
        for p in self.pipes:
            p.drain()


def bind_to_gpu_numa_node(device_index: int) -> dict:
    """Pin the calling process to the CPUs of the NUMA node its GPU hangs off, so that the pinned host
    buffers it allocates afterwards (first touch) and the threads that feed the GPU sit next to the GPU's PCIe
    root port.  With one process per GPU on a two-socket box this keeps host<->device copies off the
    inter-socket link.  Returns what was done; never raises (a box without sysfs / NVML is left alone)."""
    import os
    info = {"device": device_index, "numa_node": None, "cpus": None}
    try:
        import pynvml
        pynvml.nvmlInit()
        visible = os.environ.get("CUDA_VISIBLE_DEVICES")
        phys = device_index
        if visible:
            ids = [v.strip() for v in visible.split(",") if v.strip()]
            if device_index < len(ids) and ids[device_index].isdigit():
                phys = int(ids[device_index])
        h = pynvml.nvmlDeviceGetHandleByIndex(phys)
        bus = pynvml.nvmlDeviceGetPciInfo(h).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        dom, rest = bus.split(":", 1)
        path = f"/sys/bus/pci/devices/{dom[-4:].lower()}:{rest.lower()}/numa_node"
        node = int(open(path).read().strip())
        info["numa_node"] = node
        if node < 0:
            return info
        cpulist = open(f"/sys/devices/system/node/node{node}/cpulist").read().strip()
        cpus = set()
        for part in cpulist.split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = cpus & os.sched_getaffinity(0)
        if allowed:
            os.sched_setaffinity(0, allowed)
            info["cpus"] = len(allowed)
    except Exception as ex:  # best effort
        info["error"] = str(ex)[:120]
    return info

"""Host-side logic of the multi-GPU path on CPU: instance sharding and the (objective, violation)
all-gather over gloo with world_size 2 (the GPU run uses the same code over NCCL)."""
import os
import socket

import numpy as np
import pytest


def test_shard_range_partitions():
    from ctdirect_b200.batch import shard_range
    for n, w in ((4096, 8), (4096, 1), (10, 4), (3, 8), (0, 2)):
        parts = [shard_range(n, w, r) for r in range(w)]
        assert parts[0][0] == 0 and parts[-1][1] == n
        assert all(parts[r][1] == parts[r + 1][0] for r in range(w - 1))
        sizes = [hi - lo for lo, hi in parts]
        assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(10, 2, 2)


def _worker(rank, world, port, n, q):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from ctdirect_b200.batch import gather_norms, shard_range
        lo, hi = shard_range(n, world, rank)
        obj = torch.arange(lo, hi, dtype=torch.float64) * 1.5
        viol = torch.arange(lo, hi, dtype=torch.float64) * -2.0
        g_obj, g_viol = gather_norms(obj, viol, n)
        q.put((rank, g_obj.numpy().copy(), g_viol.numpy().copy()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n", [7, 8])
def test_gather_norms_gloo_world2(n):
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=60) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, g_obj, g_viol in res:
        np.testing.assert_array_equal(g_obj, np.arange(n) * 1.5)
        np.testing.assert_array_equal(g_viol, np.arange(n) * -2.0)


def test_gather_norms_single_process():
    import torch
    from ctdirect_b200.batch import gather_norms
    o, v = gather_norms(torch.tensor([1.0, 2.0], dtype=torch.float64), torch.tensor([0.5, 0.0], dtype=torch.float64), 2)
    assert o.tolist() == [1.0, 2.0] and v.tolist() == [0.5, 0.0]


def _pipe_worker(rank, world, port, n_local, q):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from ctdirect_b200.batch import NormsPipeline
        pipe = NormsPipeline(n_local, "cpu", depth=2)
        got = []
        for k in range(5):
            def ev(slot, obj, viol, k=k):
                obj.copy_(torch.arange(n_local, dtype=torch.float64) + 100.0 * rank + 1000.0 * k)
                viol.fill_(float(k - rank))
            s = pipe.submit(ev)
            got.append(pipe.result(s).clone().numpy())
        pipe.drain()
        q.put((rank, got))
    finally:
        dist.destroy_process_group()


def test_norms_pipeline_gloo_world2():
    """The pipelined gather (serial on CPU tensors) returns every evaluation's norms of every rank, in order."""
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    n_local = 3
    procs = [ctx.Process(target=_pipe_worker, args=(r, 2, port, n_local, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=60) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, got in res:
        for k, g in enumerate(got):
            assert g.shape == (2, 2, n_local)
            for r in range(2):
                np.testing.assert_array_equal(g[r, 0], np.arange(n_local) + 100.0 * r + 1000.0 * k)
                np.testing.assert_array_equal(g[r, 1], np.full(n_local, float(k - r)))


def test_numa_binding_helper_never_raises():
    """Best effort by contract: without NVML / sysfs (this container) it reports and leaves the process alone."""
    from ctdirect_b200.batch import bind_to_gpu_numa_node
    before = os.sched_getaffinity(0)
    info = bind_to_gpu_numa_node(0)
    assert info["device"] == 0 and "numa_node" in info
    assert os.sched_getaffinity(0) <= before and len(os.sched_getaffinity(0)) > 0


def test_native_pipeline_rejects_unequal_shards():
    from ctdirect_b200.batch import NativeNormsPipeline

    class FakeDocp:
        batch, device = 3, 0

    with pytest.raises(ValueError, match="same size on every rank"):
        NativeNormsPipeline(FakeDocp(), 0, 2, b"x" * 128, n_instances=7)
    with pytest.raises(ValueError, match="depth"):
        NativeNormsPipeline(FakeDocp(), 0, 2, b"x" * 128, depth=9)

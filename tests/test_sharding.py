"""Host-side logic of the multi-GPU path on CPU: world-size-2 gloo.  Frames and state shard by
subscriber MAC hash, there is no data-path collective; the only collective is the all-reduce of
the packed counter vector, which this test performs over gloo with per-shard ORACLE counters
standing in for the per-GPU counters (the reduction must equal the unsharded run)."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    from bng_b200 import workloads as W
    from bng_b200.layouts import as_bytes
    from oracle.pyoracle import Oracle, available
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n = 4096
    wl = W.pipeline(n, rank, world, n_subs=512, flows_per_sub=4, imix=True)
    o = Oracle("reference" if available("reference") else "port")
    for m, k, v in wl.maps:
        assert o.update_batch(m, as_bytes(k), as_bytes(v)) == 0
    for prog, h, l in wl.prewarm:
        pa = o.arena(h.shape[0] * 64 + 64)
        pa[: h.shape[0] * 64] = h.reshape(-1)
        o.run(prog, pa, l.copy(), wl.now0 - 1, stride=64)
    arena = o.arena(n * 64)
    arena[:] = wl.headers.reshape(-1)
    o.run(wl.prog, arena, wl.lens.copy(), wl.now0, stride=64)
    local = np.concatenate([o.lookup(m, np.zeros(4, np.uint8)).view("<u8") for m in
                            ("antispoof_stats", "qos_stats_map", "nat_stats_map", "stats_map")]).astype(np.int64)
    t = torch.from_numpy(local.copy())
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    subs = torch.tensor([wl.n_subs_local])
    dist.all_reduce(subs)
    if rank == 0:
        q.put((t.numpy().tolist(), int(subs.item()), local.tolist()))
    dist.destroy_process_group()


def test_two_shards_reduce_to_consistent_totals():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    total, subs, local0 = q.get(timeout=300)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert subs == 512                       # every subscriber lives on exactly one shard
    assert total[0] + total[1] == 2 * 4096   # antispoof: every frame of both shards allowed or dropped
    assert total[10] > 0 and total[10] >= local0[10]  # SNAT counters add up across shards


def test_shard_function_partitions_subscribers():
    sys.path.insert(0, ROOT)
    from bng_b200 import workloads as W
    parts = [set(W.local_subscribers(10_000, r, 8).tolist()) for r in range(8)]
    assert sum(len(p) for p in parts) == 10_000 and len(set().union(*parts)) == 10_000
    assert min(len(p) for p in parts) > 1000

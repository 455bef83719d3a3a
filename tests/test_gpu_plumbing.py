"""Product-side plumbing: staged upserts, counter reconciliation entry points, host arenas, several contexts in
one process."""
import numpy as np
import pytest

from bng_b200 import layouts as L
from bng_b200 import synth as S

pytestmark = pytest.mark.gpu


def _dp(**kw):
    from bng_b200 import Dataplane
    opts = dict(max_subscribers=1 << 14, max_nat_sessions=1 << 14, max_eim_mappings=1 << 14, max_batch=1 << 14)
    opts.update(kw)
    return Dataplane(**opts)


def test_staged_upserts_apply_at_the_batch_boundary_and_before_reads():
    n = 5000
    keys, v = S.bindings(n)
    dp, ref = _dp(), _dp()
    try:
        # every key twice, the second value differing: the LAST staged value must win (Puts applied in order)
        first = v.copy()
        first["ipv4_addr"] = 0
        for i in range(n):
            assert dp.update_staged("subscriber_bindings", keys[i], first[i]) == 0
        for i in range(n):
            assert dp.update_staged("subscriber_bindings", keys[i], v[i]) == 0
        inf = dp.staged_info()
        assert inf["pending"] == 2 * n and inf["flushes"] == 0
        # read-your-writes: a lookup of the same map flushes it first
        got = dp.lookup("subscriber_bindings", keys[17])
        assert got is not None and bytes(got) == bytes(L.as_bytes(v[17:18])[0])
        inf = dp.staged_info()
        assert inf["pending"] == 0 and inf["flushes"] == 1 and inf["errors"] == 0
        assert ref.update_batch("subscriber_bindings", keys, v) == 0
        for a, b in zip(dp.dump("subscriber_bindings"), ref.dump("subscriber_bindings")):
            assert np.array_equal(a, b)
        # staged, then a batch: visible to the program without any explicit flush
        k2, v2 = S.bindings(n + 64)
        cfg = np.zeros(1, L.antispoof_config)
        cfg["default_mode"] = 1
        for d in (dp, ref):
            assert d.update("antispoof_config", np.uint32(0), cfg) == 0
        for i in range(n, n + 64):
            dp.update_staged("subscriber_bindings", k2[i], v2[i])
        sub = np.arange(n, n + 64)
        lens = np.full(64, 64, np.uint32)
        hdr = S.ipv4_headers(S.sub_mac_key(sub), np.uint64(0x02FFFFFFFFFE), S.sub_ip(sub), np.uint32(0x08080808), 6, 4000, 443, lens)
        vd = dp.run("antispoof_ingress", hdr.reshape(-1).copy(), lens.copy(), 10**9, stride=64)
        assert (np.asarray(vd) == 0).all(), "staged bindings were not visible to the batch"
        vr = ref.run("antispoof_ingress", hdr.reshape(-1).copy(), lens.copy(), 10**9, stride=64)
        assert (np.asarray(vr) == 2).all()  # (the other context never saw them: strict mode drops unknown MACs)
        # a staged Put into a full map is an error counted at the flush
        tiny = _dp(max_subscribers=64)
        try:
            for i in range(200):
                tiny.update_staged("subscriber_bindings", keys[i], v[i])
            tiny.sync()
            assert tiny.staged_info()["errors"] == 200 - 64
        finally:
            tiny.close()
    finally:
        dp.close()
        ref.close()


def test_sync_reduce_single_rank_communicator():
    """bng_comm_init / bng_sync_reduce with a world of one (the N > 1 case runs under torchrun: bench.py)."""
    from bng_b200 import Dataplane
    dp = _dp()
    try:
        keys, v = S.bindings(100)
        assert dp.update_batch("subscriber_bindings", keys, v) == 0
        cfg = np.zeros(1, L.antispoof_config)
        cfg["default_mode"] = 1
        dp.update("antispoof_config", np.uint32(0), cfg)
        sub = np.arange(200) % 150
        lens = np.full(200, 64, np.uint32)
        hdr = S.ipv4_headers(S.sub_mac_key(sub), np.uint64(0x02FFFFFFFFFE), S.sub_ip(sub), np.uint32(0x08080808), 6, 4000, 443, lens)
        dp.run("antispoof_ingress", hdr.reshape(-1).copy(), lens, 10**9, stride=64)
        local = dp.sync_reduce()  # no communicator yet: this context's own counters
        assert np.array_equal(local[:6], dp.stats("antispoof_stats"))
        try:
            uid = Dataplane.comm_unique_id()
        except OSError:
            pytest.skip("no libnccl.so.2 resolvable in this process")
        dp.comm_init(uid, 0, 1)
        tot = dp.sync_reduce()
        assert np.array_equal(tot, local) and tot[0] + tot[1] == 200
    finally:
        dp.close()


def test_two_contexts_one_process_every_program():
    """No process-wide launch state: a second context (on the second GPU when the box has one) runs every program,
    including the DHCP kernel that needs a per-device shared-memory attribute."""
    import torch
    import harness
    import scenarios
    dev2 = 1 if torch.cuda.device_count() > 1 else 0
    for name in ("dhcp", "pipeline"):
        res = []
        for dev in (0, dev2):
            be = harness.GpuBackend(device=dev)
            try:
                res.append(harness.run_script(be, scenarios.ALL_SCRIPTS[name]()))
            finally:
                be.close()
        harness.compare(res[0], res[1], f"{name}: context on device 0 vs second context on device {dev2}")
    a, b = harness.GpuBackend(device=0), harness.GpuBackend(device=dev2)  # and two alive at the same time
    try:
        ra = harness.run_script(a, scenarios.ALL_SCRIPTS["dhcp"]())
        rb = harness.run_script(b, scenarios.ALL_SCRIPTS["dhcp"]())
        harness.compare(ra, rb, "dhcp: two live contexts")
    finally:
        a.close()
        b.close()


def test_host_arena_alloc_free_cycles():
    import ctypes
    from bng_b200.dataplane import load_library
    lib = load_library()
    for size in (1 << 12, (3 << 20) + 17, 64 << 20):
        p = lib.bng_host_alloc(size)
        assert p
        view = np.ctypeslib.as_array((ctypes.c_uint8 * size).from_address(p))
        view[:] = 7
        assert int(view[-1]) == 7
        lib.bng_host_free(p)
    lib.bng_host_free(None)


def test_snapshot_restores_into_a_context_of_another_size():
    """bng_snapshot / bng_restore (HA hand-over): the state a pipeline scenario leaves behind is carried
    into a fresh context with different table capacities; the second half of the traffic then behaves identically."""
    import harness
    import scenarios
    sc = scenarios.pipeline_script()
    runs = [i for i, st in enumerate(sc.steps) if st[0] == "run"]
    cut = runs[2]  # state after two batches moves to the other context
    first, second = harness.Script("a"), harness.Script("b")
    first.steps, second.steps = sc.steps[:cut], sc.steps[cut:]
    a = harness.GpuBackend()
    try:
        harness.run_script(a, first, tables=())
        blob = a.dp.snapshot()
        ra = harness.run_script(a, second)
    finally:
        a.close()
    b = harness.GpuBackend(max_subscribers=1 << 12, max_nat_sessions=1 << 15, max_eim_mappings=1 << 13)
    try:
        b.dp.restore(blob)
        rb = harness.run_script(b, second)
    finally:
        b.close()
    harness.compare(ra, rb, "after the snapshot: original context vs restored context")
    assert len(blob) > 10_000 and len(ra["tk_nat_sessions"]) > 100


def test_update_batch_with_repeated_keys_applies_in_order():
    """bpf(2) semantics inside one bng_map_update_batch: entries take effect one after the other — the last value of a
    repeated key stays (BPF_ANY), its second occurrence fails under BPF_NOEXIST."""
    dp = _dp()
    try:
        keys, v = S.bindings(8)
        k = np.concatenate([keys[:4], keys[:4], keys[4:]])
        vv = np.concatenate([v[:4], v[4:8], v[4:]])  # second occurrence of keys 0..3 carries other values
        assert dp.update_batch("subscriber_bindings", k, vv) == 0
        for i in range(4):
            assert bytes(dp.lookup("subscriber_bindings", keys[i])) == bytes(L.as_bytes(v[4 + i:5 + i])[0])
        assert dp.map_info("subscriber_bindings")["count"] == 8
        import errno
        assert dp.update_batch("subscriber_bindings", k, vv, 1) == -errno.EEXIST  # NOEXIST: everything exists by now
        dp.clear("subscriber_bindings")
        assert dp.update_batch("subscriber_bindings", k, vv, 1) == -errno.EEXIST  # ... and the repeats clash with themselves
        assert dp.map_info("subscriber_bindings")["count"] == 8
        for i in range(4):  # the FIRST occurrence won under NOEXIST
            assert bytes(dp.lookup("subscriber_bindings", keys[i])) == bytes(L.as_bytes(v[i:i + 1])[0])
    finally:
        dp.close()

"""Product-side plumbing: staged upserts, counter reconciliation entry points, host arenas, several contexts in
one process."""
import numpy as np
import pytest

from bng_b200 import layouts as L
from bng_b200 import synth as S
from bng_b200 import workloads as W
from bng_b200.layouts import as_bytes

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


@pytest.mark.parametrize("name,ring", [("antispoof_64", "spoof_events"), ("pipeline_64", "spoof_events"), ("nat_cold_64", "nat_log_rb")])
def test_event_staging_ring_overflow_is_counted_not_fatal(name, ring):
    """A staging ring smaller than what one batch logs: the records that found a slot come out, the rest are counted
    in bng_events_lost, and nothing else about the batch changes (verdicts, counters — packets_logged counts a
    violation whether or not its record could be written, bpf/antispoof.c:171-174)."""
    from bng_b200 import Dataplane
    n = 1 << 13 if name == "nat_cold_64" else 1 << 15  # (nat_log_rb also has the reference's 1 MiB ring rule: stay under it)
    wl = W.build(name, n)

    def run(cap):
        dp = Dataplane(max_batch=n, event_capacity=cap, **W.sizing(wl))
        try:
            for m, k, val in wl.maps:
                assert dp.update_batch(m, as_bytes(k), as_bytes(val)) == 0, m
            for prog, h, l in wl.prewarm:
                dp.run(prog, h.reshape(-1).copy(), l.copy(), wl.now0 - 1, stride=64)
                dp.drain("nat_log_rb")
            lost0 = dp.events_lost
            a = np.ascontiguousarray(wl.headers).reshape(-1).copy()
            v = dp.run(wl.prog, a, wl.lens.copy(), wl.now0, stride=wl.headers.shape[1])
            ev = dp.drain(ring)
            stats = {m: dp.stats(m).copy() for m in ("antispoof_stats", "nat_stats_map", "qos_stats_map")}
            return np.asarray(v).copy(), a, ev, dp.events_lost - lost0, stats
        finally:
            dp.close()

    v_big, a_big, ev_big, lost_big, st_big = run(1 << 18)
    assert lost_big == 0 and ev_big.shape[0] > 200, ev_big.shape
    cap = 96
    v_small, a_small, ev_small, lost_small, st_small = run(cap)
    assert ev_small.shape[0] == cap
    assert lost_small == ev_big.shape[0] - cap
    assert np.array_equal(v_big, v_small) and np.array_equal(a_big, a_small)
    for m in st_big:
        assert np.array_equal(st_big[m], st_small[m]), m
    # every record that did come out is one of the records of the unconstrained run
    big = {bytes(r) for r in ev_big}
    assert all(bytes(r) in big for r in ev_small)

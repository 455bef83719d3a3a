"""Sharded GPU runs against the UNSHARDED reference: the frames of one pipeline_imix batch are
steered to N shards by bng_shard_of_mac(source MAC) — each shard a dataplane context of its own holding only its
subscribers' state — and the union of what the shards produce must be bit-identical to one reference run over the
whole batch: per-frame verdicts and bytes, the summed counters, the union of the table dumps and of the event
streams.  (The shards run one after the other on one device: that is enough, they share nothing.)"""
import numpy as np
import pytest

from bng_b200 import layouts as L
from bng_b200 import synth as S
from bng_b200 import workloads as W
from bng_b200.layouts import as_bytes

pytestmark = pytest.mark.gpu

STATS = ("antispoof_stats", "qos_stats_map", "nat_stats_map")
TABLES = ("nat_sessions", "nat_reverse", "eim_table", "subscriber_nat", "qos_ingress", "subscriber_bindings")
KEYED_BY_SUBSCRIBER = {"subscriber_bindings": "mac", "qos_ingress": "ip", "subscriber_nat": "ip"}


def _frame_mac_keys(headers):
    k = np.zeros(len(headers), np.uint64)
    for i in range(6):
        k = (k << np.uint64(8)) | headers[:, 6 + i].astype(np.uint64)
    return k


def _ip_to_shard(n_subs, world):
    sub = np.arange(n_subs, dtype=np.uint32)
    shard = S.shard_of_mac(S.sub_mac_key(sub), world)
    return {bytes(k): int(s) for k, s in zip(S.ip_bytes(S.sub_ip(sub)), shard)}


def _oracle(wl, steps):
    from oracle.pyoracle import Oracle, available
    o = Oracle("reference" if available("reference") else "port")
    for m, k, v in wl.maps:
        assert o.update_batch(m, as_bytes(k), as_bytes(v)) == 0
    for prog, h, l in wl.prewarm:
        pa = o.arena(h.shape[0] * 64 + 64)
        pa[: h.shape[0] * 64] = h.reshape(-1)
        o.run(prog, pa, l.copy(), wl.now0 - 1, stride=64)
    for m in ("spoof_events", "nat_log_rb"):
        o.drain(m)  # the log consumer has caught up before the batches under test (ONE ring here, one per shard there)
    outs = []
    for s in range(steps):
        oa = o.arena(wl.n * 64 + 64)
        oa[: wl.n * 64] = wl.headers.reshape(-1)
        v = o.run(wl.prog, oa, wl.lens.copy(), wl.now0 + s * wl.now_step, stride=64)
        outs.append((np.asarray(v).copy(), np.array(oa[: wl.n * 64]).reshape(-1, 64)))
        o.free_arenas()
    stats = {m: o.lookup(m, np.zeros(4, np.uint8)).view("<u8").copy() for m in STATS}
    events = {m: o.drain(m) for m in ("spoof_events", "nat_log_rb")}
    tables = {m: o.dump(m) for m in TABLES}
    return outs, stats, events, tables


def _rows(a):
    return sorted(bytes(r) for r in a)


@pytest.mark.parametrize("world", [2, 8])
def test_union_of_shards_equals_unsharded_reference(world):
    from bng_b200 import Dataplane
    n, n_subs, steps = 1 << 18, 2_000, 2  # ~130 frames per subscriber and batch: the lower tiers drop
    wl = W.pipeline(n, 0, 1, n_subs=n_subs, flows_per_sub=16, imix=True)  # the whole batch, unsharded
    ref_outs, ref_stats, ref_events, ref_tables = _oracle(wl, steps)
    ip_shard = _ip_to_shard(n_subs, world)
    frame_shard = S.shard_of_mac(_frame_mac_keys(wl.headers), world)
    warm_h, warm_l = wl.prewarm[0][1], wl.prewarm[0][2]
    warm_shard = S.shard_of_mac(_frame_mac_keys(warm_h), world)
    verdicts = [np.full(n, 255, np.uint8) for _ in range(steps)]
    frames = [np.zeros((n, 64), np.uint8) for _ in range(steps)]
    stats = {m: 0 for m in STATS}
    events = {"spoof_events": [], "nat_log_rb": []}
    tables = {m: ([], []) for m in TABLES}
    for rank in range(world):
        dp = Dataplane(max_batch=n, max_subscribers=4 * n_subs // world + 1024, max_nat_sessions=1 << 19, max_eim_mappings=1 << 19,
                       rank=rank, world=world)
        try:
            for m, k, v in wl.maps:  # the control plane routes each per-subscriber upsert to the owner, replicates the rest
                kb, vb = as_bytes(k), as_bytes(v)
                how = KEYED_BY_SUBSCRIBER.get(m)
                if how == "mac":
                    keep = S.shard_of_mac(k.astype(np.uint64), world) == rank
                    kb, vb = kb[keep], vb[keep]
                elif how == "ip":
                    keep = np.array([ip_shard[bytes(x)] == rank for x in kb])
                    kb, vb = kb[keep], vb[keep]
                assert dp.update_batch(m, kb, vb) == 0, m
            mine_w = warm_shard == rank
            dp.run("nat44_egress", warm_h[mine_w].reshape(-1).copy(), warm_l[mine_w].copy(), wl.now0 - 1, stride=64)
            for m in ("spoof_events", "nat_log_rb"):
                dp.drain(m)
            mine = np.nonzero(frame_shard == rank)[0]  # index order is kept inside a shard
            for s in range(steps):
                a = wl.headers[mine].reshape(-1).copy()
                v = dp.run(wl.prog, a, wl.lens[mine].copy(), wl.now0 + s * wl.now_step, stride=64)
                verdicts[s][mine] = v
                frames[s][mine] = a.reshape(-1, 64)
            for m in STATS:
                stats[m] = stats[m] + dp.stats(m).astype(np.uint64)
            for m in events:
                events[m].append(dp.drain(m))
            for m in TABLES:
                k, v = dp.dump(m)
                tables[m][0].append(k)
                tables[m][1].append(v)
            assert dp.lru_overflow == 0 and dp.events_lost == 0
        finally:
            dp.close()
    for s in range(steps):
        assert np.array_equal(verdicts[s], ref_outs[s][0]), f"step {s}: verdicts differ"
        assert np.array_equal(frames[s], ref_outs[s][1]), f"step {s}: frame bytes differ"
    for m in STATS:
        assert np.array_equal(stats[m], ref_stats[m]), f"{m}: {stats[m]} vs {ref_stats[m]}"
    for m, parts in events.items():
        got = np.concatenate([p for p in parts if len(p)], axis=0) if any(len(p) for p in parts) else np.zeros((0, 1), np.uint8)
        w = got.shape[1] - (4 if m == "nat_log_rb" else 0)  # trailing padding of nat_log_entry
        assert _rows(got[:, :w]) == _rows(ref_events[m][:, :w]), f"{m}: event multisets differ"
    for m in TABLES:
        k = np.concatenate(tables[m][0], axis=0)
        v = np.concatenate(tables[m][1], axis=0)
        rk, rv = ref_tables[m]
        for off, ln in L.PADDING.get(m, ()):
            v[:, off:off + ln] = 0
            rv = rv.copy()
            rv[:, off:off + ln] = 0
        assert _rows(np.concatenate([k, v], axis=1)) == _rows(np.concatenate([rk, rv], axis=1)), f"{m}: table contents differ"

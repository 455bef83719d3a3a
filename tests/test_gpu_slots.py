"""Frames whose length claims more than their slot holds (a header-split receive ring: 64-byte slots, len = the full
frame length).  The programs' bounds checks run against the bytes PRESENT (data_end - data in the reference), byte
counters against len (skb->len): a frame whose L4 header lies beyond its slot (IPv4 options) is passed untouched and
its neighbours are never read or written; every other frame is processed exactly as in a whole-frame arena.
ADVICE r1 (high): the zero-copy path used to let such a frame rewrite the next frame's MAC / address bytes."""
import numpy as np
import pytest

import scenarios
from bng_b200 import synth as S
from harness import Script

pytestmark = pytest.mark.gpu
GW = scenarios.GW_MAC


def _frames(n=6000, n_subs=20, seed=0x51A7):
    r = scenarios.rng(seed)
    longopt = r.integers(0, 8, n) == 0
    # subscribers 0..9 send plain frames, 10..19 frames with long options: the two sets share no NAT / QoS state
    sub = np.where(longopt, 10 + r.integers(0, 10, n), r.integers(0, 10, n))
    proto = r.choice(np.array([6, 17, 1], dtype=np.uint32), n)
    lens = r.choice(np.array([64, 594, 1518], dtype=np.uint32), n)
    lens = np.where(longopt, np.maximum(lens, 594), lens).astype(np.uint32)
    h = S.ipv4_headers(S.sub_mac_key(sub), np.uint64(GW), S.sub_ip(sub), np.uint32(0x08080808), proto,
                       (30000 + r.integers(0, 4, n)).astype(np.uint32), 443, lens, l4_check=0x1357)
    ihl = np.where(proto == 6, r.integers(9, 16, n), r.integers(12, 16, n))  # L4 header ends past byte 63
    h[longopt, 14] = (0x40 | ihl[longopt]).astype(np.uint8)
    h[longopt, 34:64] = r.integers(0, 256, (int(longopt.sum()), 30), dtype=np.uint8)
    return h, lens, longopt, sub


def _maps(dp, n_subs=20):
    sc = Script("maps")
    k, v = S.bindings(n_subs)
    sc.update("subscriber_bindings", k, v)
    scenarios.nat_maps(sc, n_subs, 64, 0x0F)
    qk, qv = S.qos_buckets(n_subs)
    qv["burst_bytes"] = np.minimum(qv["burst_bytes"], 40000)
    qv["tokens"] = qv["burst_bytes"]
    sc.update("qos_ingress", qk, qv)
    for st in sc.steps:
        assert dp.update_batch(st[1], st[2], st[3], st[4]) == 0, st[1]


def _run(prog, arena, lens, stride, mode):
    import torch
    from bng_b200 import MEM_DEVICE, MEM_HOST, Dataplane
    dp = Dataplane(max_subscribers=1 << 10, max_nat_sessions=1 << 14, max_eim_mappings=1 << 14, max_batch=1 << 14)
    try:
        _maps(dp)
        if mode == "device":
            ta, tl = torch.from_numpy(arena.copy()).cuda(), torch.from_numpy(lens.view(np.int32).copy()).cuda()
            tv = torch.zeros(len(lens), dtype=torch.uint8, device="cuda")
            torch.cuda.synchronize()
            dp.run(prog, ta, tl, 10**9, stride=stride, verdict=tv, mem=MEM_DEVICE)
            dp.sync()
            a, v = ta.cpu().numpy(), tv.cpu().numpy()
        elif mode == "pinned":
            ta, tl = torch.from_numpy(arena.copy()).pin_memory(), torch.from_numpy(lens.view(np.int32).copy()).pin_memory()
            tv = torch.zeros(len(lens), dtype=torch.uint8).pin_memory()
            dp.run(prog, ta, tl, 10**9, stride=stride, verdict=tv, mem=MEM_HOST, arena_bytes=arena.nbytes)
            a, v = ta.numpy().copy(), tv.numpy().copy()
        else:
            a = arena.copy()
            v = dp.run(prog, a, lens.copy(), 10**9, stride=stride)
        st = {m: dp.stats(m) for m in ("nat_stats_map", "qos_stats_map", "antispoof_stats")}
        assert dp.lru_overflow == 0
        return a.reshape(-1, stride), np.asarray(v), st
    finally:
        dp.close()


@pytest.mark.parametrize("mode", ["device", "pageable", "pinned"])
@pytest.mark.parametrize("prog", ["nat44_egress", "pipeline_up"])
def test_header_ring_slots_are_isolated(prog, mode):
    h, lens, longopt, sub = _frames()
    ring = scenarios.fixed(h, 64)     # 64-byte slots, len[] = full frame length
    whole = scenarios.fixed(h, 1536)  # every frame in a slot that holds all of it
    a64, v64, st64 = _run(prog, ring, lens, 64, mode)
    a1536, v1536, st1536 = _run(prog, whole, lens, 1536, "device")
    # frames whose L4 header is not in the slot: passed, byte-identical — and so are their neighbours' other bytes
    assert np.array_equal(a64[longopt], h[longopt]), "a frame with options beyond its slot was modified"
    if prog == "nat44_egress":
        assert (v64[longopt] == 0).all()  # (in the pipeline the token bucket may still drop them)
    # every other frame: exactly what the whole-frame arena gives (verdicts and the 64 header bytes)
    plain = ~longopt
    assert np.array_equal(v64[plain], v1536[plain])
    assert np.array_equal(a64[plain], a1536[plain][:, :64])
    # ... and the long-option frames WERE translated when all their bytes were there (the test is not vacuous)
    assert (a1536[longopt][:, 26:30] != h[longopt][:, 26:30]).any()
    assert st64["nat_stats_map"][0] < st1536["nat_stats_map"][0]

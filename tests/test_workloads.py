"""The synthetic workloads bench.py measures (bng_b200/workloads.py), checked on the CPU: shapes, sharding, and —
through the oracle — that each workload exercises the path it is named after (hits where it promises hits, misses
where it promises misses)."""
import numpy as np
import pytest

from bng_b200 import layouts as L
from bng_b200 import workloads as W
from bng_b200.layouts import as_bytes
from oracle import pyoracle

KIND = "reference" if pyoracle.available("reference") else ("port" if pyoracle.available("port") else None)
needs_oracle = pytest.mark.skipif(KIND is None, reason="no oracle library built")


def run_workload(wl, steps=1):
    """maps -> prewarm -> derive -> `steps` passes; returns (oracle, verdicts of the last pass)."""
    o = pyoracle.Oracle(KIND)
    for m, k, v in wl.maps:
        o.update_batch(m, as_bytes(k), as_bytes(v))
    translated = []
    for prog, h, l in wl.prewarm:
        pa = o.arena(h.shape[0] * 64 + 64)
        pa[: h.shape[0] * 64] = h.reshape(-1)
        o.run(prog, pa, l.copy(), wl.now0 - 1, stride=64)
        translated.append(np.array(pa[: h.shape[0] * 64]))
    if wl.derive is not None:
        wl.headers, wl.lens = wl.derive(translated)
    off16, stride, total16 = W.slot16(wl.lens, wl.imix, wl.headers.shape[1])
    hw = wl.headers.shape[1]
    arena = o.arena(total16 * 16 + 64)
    v = None
    for s in range(steps):
        if off16 is None:
            arena[: wl.n * stride].reshape(wl.n, stride)[:, :hw] = wl.headers
        else:
            a16 = arena[: total16 * 16].reshape(total16, 16)
            for g in range(hw // 16):
                a16[off16.astype(np.int64) + g] = wl.headers[:, 16 * g: 16 * g + 16]
        v = o.run(wl.prog, arena, wl.lens.copy(), wl.now0 + s * wl.now_step, off16=off16, stride=stride)
    return o, v


def stats(o, name, dtype):
    return np.frombuffer(bytes(o.lookup(name, np.zeros(4, np.uint8))), dtype)[0]


def test_every_builder_produces_consistent_shapes():
    for name, build in W.BUILDERS.items():
        wl = build(2048, 0, 1) if name != "dhcp" else W.dhcp(2048, 0, 1, n_subs=4096)
        assert wl.name == name
        assert wl.headers.shape[0] == wl.lens.shape[0] == wl.n
        assert wl.headers.dtype == np.uint8 and wl.headers.shape[1] % 16 == 0
        assert name in W.ALGO_BYTES
        cap = W.sizing(wl)
        assert cap["max_subscribers"] >= wl.n_subs_local


def test_sharding_partitions_the_subscribers():
    whole = W.local_subscribers(5000, 0, 1)
    parts = [W.local_subscribers(5000, r, 4) for r in range(4)]
    assert sum(len(p) for p in parts) == len(whole) == 5000
    assert len(np.unique(np.concatenate(parts))) == 5000
    assert all(len(p) > 1000 for p in parts)  # splitmix64 spreads them


def test_imix_layout_is_64_byte_aligned_and_disjoint():
    wl = W.pipeline(4096, 0, 1, n_subs=64, flows_per_sub=4, imix=True)
    off16, stride, total16 = W.slot16(wl.lens, True, 64, 64)
    assert stride == 0 and (off16 % 4 == 0).all()  # 64-byte boundaries
    end = off16.astype(np.int64) * 16 + wl.lens
    assert (end[:-1] <= off16[1:].astype(np.int64) * 16).all() and end[-1] <= total16 * 16
    assert abs(float(wl.lens.mean()) - 361.8) < 25  # IMIX 7:4:1 of 64/594/1518


@needs_oracle
def test_nat_steady_hits_and_cold_misses():
    o, v = run_workload(W.nat(4096, 0, 1, n_subs=64, flows_per_sub=8, cold=False))
    st = stats(o, "nat_stats_map", L.nat_stats)
    assert st["sessions_created"] == 64 * 8 and st["packets_snat"] == 64 * 8 + 4096 and (v == 0).all()
    o, v = run_workload(W.nat(10 ** 6, 0, 1, n_subs=64, flows_per_sub=8, cold=True))
    st = stats(o, "nat_stats_map", L.nat_stats)
    assert st["sessions_created"] == 64 * 8 == st["packets_snat"] and st["eim_misses"] == 64 * 8


@needs_oracle
def test_nat_ingress_workload_is_all_dnat_hits():
    """The replies are derived from what the egress prewarm made of each flow: every one must find its reverse entry."""
    n = 5000
    o, v = run_workload(W.nat_ingress(n, 0, 1, n_subs=64, flows_per_sub=8), steps=2)
    st = stats(o, "nat_stats_map", L.nat_stats)
    assert st["packets_dnat"] == 2 * n and st["packets_passed"] == 0 and st["sessions_expired"] == 0
    assert (v == 0).all()


@needs_oracle
def test_pipeline_workload_mixes_pass_spoof_and_rate_drops():
    o, v = run_workload(W.pipeline(8192, 0, 1, n_subs=64, flows_per_sub=4, imix=True), steps=2)
    a = stats(o, "antispoof_stats", L.antispoof_stats)
    q = stats(o, "qos_stats_map", L.qos_stats)
    n = stats(o, "nat_stats_map", L.nat_stats)
    assert 0 < a["packets_dropped"] < 0.03 * 2 * 8192          # ~1 % spoofed sources
    assert a["packets_allowed"] + a["packets_dropped"] == 2 * 8192
    assert n["sessions_created"] == 64 * 4                       # only the prewarm creates flows
    assert q["packets_passed"] + q["packets_dropped"] == a["packets_allowed"]
    # with enough traffic per subscriber some tiers run out of tokens: the ordered walk has work to do
    o, v = run_workload(W.pipeline(1 << 16, 0, 1, n_subs=16, flows_per_sub=4, imix=True))
    q = stats(o, "qos_stats_map", L.qos_stats)
    assert q["packets_dropped"] > 0 and q["packets_passed"] > 0


@needs_oracle
def test_qos_workloads_both_directions():
    for egress in (False, True):
        o, v = run_workload(W.qos(4096, 0, 1, n_subs=64, egress=egress))
        q = stats(o, "qos_stats_map", L.qos_stats)
        assert q["packets_passed"] + q["packets_dropped"] == 4096 and q["packets_passed"] > 0


@needs_oracle
def test_dhcp_workload_hit_rate():
    o, v = run_workload(W.dhcp(4096, 0, 1, n_subs=8192))
    d = stats(o, "stats_map", L.dhcp_stats)
    assert d["total_requests"] == 4096
    assert 0.97 < d["fastpath_hits"] / 4096 < 1.0 and d["fastpath_hits"] + d["fastpath_misses"] == 4096
    assert (v == 3).sum() == d["fastpath_hits"]  # XDP_TX

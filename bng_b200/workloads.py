"""The synthetic workloads of BASELINE.json / SURVEY.md §8(d), built per shard.

Each builder returns a :class:`Workload`: the map contents to load (numpy,
reference layouts), the frame headers (u8[n,64] or wider for DHCP), frame
lengths, and the program to run.  Subscribers are global; a shard keeps those
with ``splitmix64(mac_key) % world == rank`` and draws its frames from them,
so every subscriber's state lives on exactly one GPU and no data-path
collective is needed.
"""
from __future__ import annotations

import os
from dataclasses import dataclass, field

import numpy as np

from . import layouts as L
from . import synth as S

GW_MAC = 0x02FFFFFFFFFE

# algorithmic HBM bytes per frame (SURVEY.md §8d / BASELINE.md §3)
ALGO_BYTES = {
    "antispoof_64": 97,
    "nat_steady_64": 317,
    "nat_cold_64": 541,
    "pipeline_imix": 405,
    "pipeline_64": 405,
    "qos_64": 64 + 36 + 16 + 1,
    "dhcp": 748,
    # downstream direction, counted the way §8d counts the upstream programs:
    "nat_ingress_64": 64 + 64 + (16 + 16) + (16 + 80) + 24 + 1,  # frame r/w, nat_reverse, nat_sessions, write-back, verdict
    "qos_egress_64": 64 + 36 + 16 + 1 + 4,                       # as qos_64 plus the priority mark
}


@dataclass
class Workload:
    name: str
    prog: str
    maps: list = field(default_factory=list)  # [(map name, keys, values)]
    headers: np.ndarray = None                # u8[n, w]
    lens: np.ndarray = None                   # u32[n]
    imix: bool = False                        # frames laid out back to back at 16 B granularity
    now0: int = 1_000_000_000
    now_step: int = 1_000_000                 # bpf_ktime_get_ns() advance per batch
    prewarm: list = field(default_factory=list)  # [(prog, headers, lens)] run once before timing
    n_subs_local: int = 0
    info: dict = field(default_factory=dict)
    derive: object = None  # callable(list of prewarm arenas after their programs ran) -> (headers, lens)

    @property
    def n(self):
        return int(self.lens.shape[0])


def local_subscribers(n_subs: int, rank: int, world: int) -> np.ndarray:
    idx = np.arange(n_subs, dtype=np.uint32)
    if world <= 1:
        return idx
    return idx[S.shard_of_mac(S.sub_mac_key(idx), world) == rank]


def _pick(seed, n, hi):
    return (S.splitmix64_array(seed, n) % np.uint64(hi)).astype(np.int64)


def antispoof(n: int, rank=0, world=1, n_subs=10_000, seed=0xB2000002) -> Workload:
    """Config #2: 10k bindings (STRICT), 64 B frames: 97 % legitimate, 1 % spoofed source (drop + event),
    1 % unknown MAC (drop), 0.5 % IPv6 without binding (drop in strict), 0.5 % ARP (allow)."""
    subs = local_subscribers(n_subs, rank, world)
    keys, v = S.bindings(n_subs)
    w = Workload("antispoof_64", "antispoof_ingress")
    cfg = np.zeros(1, L.antispoof_config)
    cfg["default_mode"], cfg["log_violations"] = L.ANTISPOOF_STRICT, 0 if os.environ.get("BNG_EXP_AS_NOLOG") else 1  # (experiment knob)
    w.maps = [("subscriber_bindings", keys[subs], v[subs]), ("antispoof_config", np.zeros(1, "<u4"), cfg)]
    sub = subs[_pick(seed + rank, n, len(subs))]
    r = (S.splitmix64_array(seed ^ 0x77 + rank, n) % np.uint64(1000)).astype(np.int64)
    src = S.sub_ip(sub)
    src = np.where(r < 10, src ^ np.uint32(0x00100000), src)
    mac = S.sub_mac_key(sub)
    mac = np.where((r >= 10) & (r < 20), mac | np.uint64(0x00F000000000), mac)
    lens = np.full(n, 64, np.uint32)
    hdr = S.ipv4_headers(mac, np.uint64(GW_MAC), src, np.uint32(0x08080808), 6, 40000, 443, lens, l4_check=0x1234)
    v6 = (r >= 20) & (r < 25)
    hdr[v6, 12], hdr[v6, 13] = 0x86, 0xDD
    arp = (r >= 25) & (r < 30)
    hdr[arp, 12], hdr[arp, 13] = 0x08, 0x06
    w.headers, w.lens, w.n_subs_local = hdr, lens, len(subs)
    return w


def _nat_maps(subs_global: int, subs: np.ndarray, flags=0x0F):
    keys, v, pubs = S.nat_blocks(subs_global)
    return [("subscriber_nat", keys[subs], v[subs]),
            ("nat_config_map", np.zeros(1, "<u4"), S.nat_config(flags)),
            ("hairpin_ips", S.ip_bytes(pubs[:1000]), np.ones(len(pubs[:1000]), np.uint8)),  # the map holds 1000 (bpf/nat44.c:41)
            ("alg_ports", np.array([(21 << 16) | 6], "<u4"), np.array([(21, 6, 1, 0)], L.alg_config))]


def nat(n: int, rank=0, world=1, n_subs=16_384, flows_per_sub=64, cold=False, seed=0xB2000003) -> Workload:
    """Config #3: 16 384 subscribers x 64 flows = 1 M 5-tuples (61 % TCP, 37 % UDP incl. a zero-checksum slice,
    2 % ICMP echo), blocks of 1024 ports via the AllocateNAT rule, flags EIM+EIF+HAIRPIN+ALG_FTP.
    cold: every flow once (100 % miss: session + reverse + EIM inserts, port allocation, log);
    steady: frames uniform over the pre-created flows (100 % hit)."""
    subs = local_subscribers(n_subs, rank, world)
    fl = S.flows(n_subs, flows_per_sub, seed)
    mine = np.nonzero(np.isin(fl["sub"], subs))[0]
    w = Workload("nat_cold_64" if cold else "nat_steady_64", "nat44_egress")
    w.maps = _nat_maps(n_subs, subs)
    every = mine
    if cold:
        pick = every[:n] if n < len(every) else every
        w.headers = S.flow_frames(fl, pick, np.full(len(pick), 64, np.uint32), udp_zero_every=97)
        w.lens = np.full(len(pick), 64, np.uint32)
    else:
        w.prewarm = [("nat44_egress", S.flow_frames(fl, every, np.full(len(every), 64, np.uint32), udp_zero_every=97),
                      np.full(len(every), 64, np.uint32))]
        pick = every[_pick(seed + 17 + rank, n, len(every))]
        w.headers = S.flow_frames(fl, pick, np.full(n, 64, np.uint32), udp_zero_every=97)
        w.lens = np.full(n, 64, np.uint32)
    w.n_subs_local = len(subs)
    w.info = {"flows": int(len(every))}
    return w


def nat_ingress(n: int, rank=0, world=1, n_subs=16_384, flows_per_sub=64, seed=0xB2000003) -> Workload:
    """SURVEY.md §8(f)-1, downstream half of config #3: return traffic of the 1 M pre-created flows (DNAT: reverse
    lookup, session lookup, TCP state, rewrite).  The replies are derived from what nat44_egress made of each
    flow's first frame: addresses and ports swapped, so the destination is the flow's public (address, port)."""
    w = nat(n, rank, world, n_subs, flows_per_sub, cold=False, seed=seed)
    w.name, w.prog = "nat_ingress_64", "nat44_ingress"
    n_flows = w.prewarm[0][1].shape[0]
    pick = _pick(seed + 19 + rank, n, n_flows)

    def derive(translated):
        t = translated[0].reshape(-1, 64)
        out = t.copy()
        out[:, 0:6], out[:, 6:12] = t[:, 6:12], t[:, 0:6]
        out[:, 26:30], out[:, 30:34] = t[:, 30:34], t[:, 26:30]
        l4 = (t[:, 23] == 6) | (t[:, 23] == 17)
        out[l4, 34:36], out[l4, 36:38] = t[l4, 36:38], t[l4, 34:36]
        out[t[:, 23] == 6, 47] = 0x10  # ACK: NEW -> ESTABLISHED on the first reply, then steady
        out[t[:, 23] == 1, 34] = 0     # echo reply
        return out[pick], np.full(n, 64, np.uint32)

    w.derive = derive
    w.headers = np.zeros((n, 64), np.uint8)  # placeholder until derive() runs (after the prewarm)
    w.lens = np.full(n, 64, np.uint32)
    return w


def pipeline(n: int, rank=0, world=1, n_subs=10_000, flows_per_sub=64, imix=True, seed=0xB2000004) -> Workload:
    """Config #4: antispoof -> NAT44 -> QoS over 10 k subscribers (binding + port block + upload bucket with
    the reference's policy tiers round-robin), 64 pre-warmed flows each, IMIX 7:4:1 or all-64 B frames,
    1 % spoofed sources; bpf_ktime_get_ns() advances 1 ms per batch so buckets refill and most tiers drop."""
    subs = local_subscribers(n_subs, rank, world)
    fl = S.flows(n_subs, flows_per_sub, seed)
    mine = np.nonzero(np.isin(fl["sub"], subs))[0]
    w = Workload("pipeline_imix" if imix else "pipeline_64", "pipeline_up", imix=imix)
    bk, bv = S.bindings(n_subs)
    qk, qv = S.qos_buckets(n_subs, upload=True)
    cfg = np.zeros(1, L.antispoof_config)
    cfg["default_mode"], cfg["log_violations"] = L.ANTISPOOF_STRICT, 1
    w.maps = [("subscriber_bindings", bk[subs], bv[subs]), ("antispoof_config", np.zeros(1, "<u4"), cfg),
              ("qos_ingress", qk[subs], qv[subs])] + _nat_maps(n_subs, subs)
    w.prewarm = [("nat44_egress", S.flow_frames(fl, mine, np.full(len(mine), 64, np.uint32), udp_zero_every=97),
                  np.full(len(mine), 64, np.uint32))]
    pick = mine[_pick(seed + 29 + rank, n, len(mine))]
    lens = S.imix_lengths(n, seed + 31 + rank) if imix else np.full(n, 64, np.uint32)
    hdr = S.flow_frames(fl, pick, lens, udp_zero_every=97)
    spoof = (S.splitmix64_array(seed + 37 + rank, n) % np.uint64(100)) == 0
    hdr[spoof, 27] ^= 0x10  # source address outside the binding
    w.headers, w.lens, w.n_subs_local = hdr, lens, len(subs)
    w.info = {"flows": int(len(mine)), "avg_frame_bytes": float(lens.mean())}
    return w


def qos(n: int, rank=0, world=1, n_subs=10_000, seed=0xB2000006, egress=False) -> Workload:
    """qos_ingress_prog keyed on the source address (upload) or — egress=True, SURVEY.md §8(f)-1 — qos_egress_prog
    keyed on the destination address (download buckets, priority mark)."""
    subs = local_subscribers(n_subs, rank, world)
    qk, qv = S.qos_buckets(n_subs, upload=not egress)
    w = Workload("qos_egress_64" if egress else "qos_64", "qos_egress_prog" if egress else "qos_ingress_prog")
    w.maps = [("qos_egress" if egress else "qos_ingress", qk[subs], qv[subs])]
    sub = subs[_pick(seed + rank, n, len(subs))]
    lens = np.full(n, 64, np.uint32)
    if egress:
        w.headers = S.ipv4_headers(np.uint64(GW_MAC), S.sub_mac_key(sub), np.uint32(0x08080808), S.sub_ip(sub), 17, 53, 5000, lens)
    else:
        w.headers = S.ipv4_headers(S.sub_mac_key(sub), np.uint64(GW_MAC), S.sub_ip(sub), np.uint32(0x08080808), 17, 5000, 53, lens)
    w.lens, w.n_subs_local = lens, len(subs)
    return w


def dhcp(n: int, rank=0, world=1, n_subs=1_000_000, seed=0xB2000005, frame_len=362) -> Workload:
    """Config #5: 1 M subscriber_pools entries over 64 ip_pools (/18, both DNS set => 50 B of options),
    362 B requests (BOOTP 320 B), 80 % REQUEST / 20 % DISCOVER, option 53 first; 99 % known MAC (XDP_TX),
    1 % unknown (XDP_PASS)."""
    subs = local_subscribers(n_subs, rank, world)
    w = Workload("dhcp", "dhcp_fastpath_prog")
    pa = np.zeros(len(subs), L.pool_assignment)
    pa["pool_id"] = 1 + (subs % 64)
    pa["allocated_ip"] = S.ip_bytes(np.uint32(0x0A000000) + subs)
    pa["lease_expiry"] = 1 << 62
    pa["client_class"] = 1
    pools = np.zeros(64, L.ip_pool)
    base = (np.uint32(0x0A000000) + (np.arange(64, dtype=np.uint32) << 14)).astype(np.uint32)
    pools["network"] = S.ip_bytes(base)
    pools["prefix_len"] = 18
    pools["gateway"] = S.ip_bytes(base + 1)
    pools["dns_primary"] = S.ip_bytes(np.full(64, 0x08080808, np.uint32))
    pools["dns_secondary"] = S.ip_bytes(np.full(64, 0x08080404, np.uint32))
    pools["lease_time"] = 3600
    cfg = np.zeros(1, L.dhcp_server_config)
    cfg["server_mac"] = [[0x02, 0xAA, 0xBB, 0xCC, 0xDD, 0x01]]
    cfg["server_ip"] = [[10, 255, 0, 1]]
    w.maps = [("subscriber_pools", S.sub_mac_key(subs), pa), ("ip_pools", np.arange(1, 65, dtype="<u4"), pools),
              ("server_config", np.zeros(1, "<u4"), cfg)]
    sub = subs[_pick(seed + rank, n, len(subs))]
    r = (S.splitmix64_array(seed ^ 0x99 + rank, n) % np.uint64(100)).astype(np.int64)
    mac = S.sub_mac_key(sub)
    mac = np.where(r == 0, mac | np.uint64(0x00F000000000), mac)
    width = ((frame_len + 15) // 16) * 16
    h = np.zeros((n, width), np.uint8)
    h[:, 0:6] = 0xFF
    h[:, 6:12] = S.mac_bytes(mac)
    h[:, 12] = 0x08
    h[:, 14] = 0x45
    h[:, 16:18] = S.port_bytes(np.full(n, frame_len - 14))
    h[:, 22], h[:, 23] = 64, 17
    h[:, 30:34] = 0xFF
    h[:, 24:26] = S.ip_checksum(h[:, 14:34])
    h[:, 34:36] = S.port_bytes(np.full(n, 68))
    h[:, 36:38] = S.port_bytes(np.full(n, 67))
    h[:, 38:40] = S.port_bytes(np.full(n, frame_len - 34))
    h[:, 42:46] = [1, 1, 6, 0]
    xid = S.splitmix64_array(seed + 5 + rank, n)
    for k in range(4):
        h[:, 46 + k] = ((xid >> np.uint64(8 * k)) & np.uint64(0xFF)).astype(np.uint8)
    h[:, 70:76] = S.mac_bytes(mac)
    h[:, 278:282] = [0x63, 0x82, 0x53, 0x63]
    h[:, 282:285] = np.stack([np.full(n, 53), np.full(n, 1), np.where(r % 5 == 1, 1, 3)], axis=1)
    h[:, 285:292] = [55, 4, 1, 3, 15, 6, 255]
    w.headers, w.lens, w.n_subs_local = h, np.full(n, frame_len, np.uint32), len(subs)
    return w


_FPS = int(os.environ.get("BNG_FLOWS_PER_SUB", "64"))  # experiment knob: flow-table footprint vs L2 size

BUILDERS = {
    "pipeline_imix": lambda n, r, wd: pipeline(n, r, wd, flows_per_sub=_FPS, imix=True),
    "pipeline_64": lambda n, r, wd: pipeline(n, r, wd, flows_per_sub=_FPS, imix=False),
    "antispoof_64": antispoof,
    "nat_steady_64": lambda n, r, wd: nat(n, r, wd, cold=False),
    "nat_cold_64": lambda n, r, wd: nat(n, r, wd, cold=True),
    "nat_ingress_64": nat_ingress,
    "qos_64": qos,
    "qos_egress_64": lambda n, r, wd: qos(n, r, wd, egress=True),
    "dhcp": dhcp,
}


def build(name: str, n: int, rank: int = 0, world: int = 1, subs_scale: int = 1) -> "Workload":
    """BUILDERS[name] with the subscriber population multiplied by subs_scale: subs_scale = world keeps the
    per-GPU population constant as GPUs are added (10 k subscribers PER GPU) instead of splitting BASELINE's 10 k
    over them — per-GPU tables then stay the same size, so a 1 -> 8 curve is not flattered by tables that shrink
    into L2."""
    if subs_scale == 1:
        return BUILDERS[name](n, rank, world)
    k = subs_scale
    if name in ("pipeline_imix", "pipeline_64"):
        return pipeline(n, rank, world, n_subs=10_000 * k, flows_per_sub=_FPS, imix=name == "pipeline_imix")
    if name == "antispoof_64":
        return antispoof(n, rank, world, n_subs=10_000 * k)
    if name in ("nat_steady_64", "nat_cold_64"):
        return nat(n, rank, world, n_subs=16_384 * k, cold=name == "nat_cold_64")
    if name == "nat_ingress_64":
        return nat_ingress(n, rank, world, n_subs=16_384 * k)
    if name in ("qos_64", "qos_egress_64"):
        return qos(n, rank, world, n_subs=10_000 * k, egress=name == "qos_egress_64")
    if name == "dhcp":
        return dhcp(n, rank, world, n_subs=1_000_000 * k)
    raise KeyError(name)


def sizing(wl: "Workload") -> dict:
    """bng_open capacities for this workload: the control plane sizes the tables for the subscribers and
    flows it provisions (2x head-room) instead of the reference's compile-time maxima."""
    headroom = int(os.environ.get("BNG_SUBS_HEADROOM", "2"))
    subs = max(1024, headroom * wl.n_subs_local)  # sparse subscriber tables: a second probe costs a memory round trip
    flows = max(4096, 2 * int(wl.info.get("flows", 0)))
    return {"max_subscribers": subs, "max_nat_sessions": flows, "max_eim_mappings": flows}


def slot16(lens: np.ndarray, imix: bool, width: int, align: int = 64):
    """Arena layout: (off16 u32[n] or None, stride, total 16-byte granules).  Variable-length frames are
    placed on `align`-byte boundaries (64 = one DRAM access granule per header, as a NIC's cache-line
    aligned receive buffers would be; 16 = densest packing the ABI allows)."""
    if not imix:
        stride = ((width + 15) // 16) * 16
        return None, stride, lens.shape[0] * stride // 16
    a16 = max(1, align // 16)
    slots = ((lens.astype(np.uint64) + align - 1) // align) * a16
    off = np.zeros(lens.shape[0], np.uint64)
    np.cumsum(slots[:-1], out=off[1:])
    total = int(off[-1] + slots[-1]) + 4
    assert total < (1 << 32)
    return off.astype(np.uint32), 0, total

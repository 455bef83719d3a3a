"""Deterministic test scripts covering the behaviours (and quirks) of the four
reference programs listed in SURVEY.md §8a.  Every script is built from seeded
PRNG data only, so the same script can be replayed against the reference
oracle, the port oracle, the GPU dataplane and the committed golden results.
"""
from __future__ import annotations

import numpy as np

from bng_b200 import layouts as L
from bng_b200 import synth as S
from harness import Script

GW_MAC = 0x02FFFFFFFFFE


def rng(seed):
    return np.random.Generator(np.random.PCG64(seed))


def fixed(frames: np.ndarray, stride: int = 64) -> np.ndarray:
    """u8[n,w] -> flat arena with `stride` bytes per frame."""
    n, w = frames.shape
    a = np.zeros((n, stride), dtype=np.uint8)
    a[:, :min(w, stride)] = frames[:, :stride]
    return a.reshape(-1)


# ---------------------------------------------------------------------------
# antispoof_ingress
# ---------------------------------------------------------------------------
def ipv6_headers(src_mac, src6: np.ndarray, n: int) -> np.ndarray:
    h = np.zeros((n, 64), dtype=np.uint8)
    h[:, 0:6] = S.mac_bytes(np.uint64(GW_MAC))
    h[:, 6:12] = S.mac_bytes(src_mac)
    h[:, 12] = 0x86
    h[:, 13] = 0xDD
    h[:, 14] = 0x60
    h[:, 20] = 6
    h[:, 21] = 64
    h[:, 22:38] = src6
    h[:, 38:54] = 0x20
    return h


def antispoof_script(seed=0xA5, n_subs=48, n=3000) -> Script:
    r = rng(seed)
    sc = Script("antispoof")
    keys, v = S.bindings(n_subs)
    v["mode"] = np.array([1, 3, 2, 0, 1, 1, 7], dtype=np.uint8)[np.arange(n_subs) % 7]
    v6 = (np.arange(n_subs) % 3) == 0
    v["ipv6_valid"] = v6
    v["ipv6_addr"][:, 0] = 0x20
    v["ipv6_addr"][:, 1] = 0x01
    v["ipv6_addr"][:, 15] = np.arange(n_subs)
    v["ipv4_valid"] = (np.arange(n_subs) % 5) != 4
    sc.update("subscriber_bindings", keys, v)
    cfg = np.zeros(1, L.antispoof_config)
    cfg["default_mode"], cfg["log_violations"] = 1, 1
    sc.update1("antispoof_config", np.uint32(0), cfg)
    # loose-mode ranges: one in wire byte order (matches 100.64/10 sources), one as the Go manager
    # would marshal it (numeric value stored little-endian), one /0-like tiny prefix on another net
    rk = np.zeros(3, L.lpm_key_v4)
    rk["prefixlen"] = [10, 10, 24]
    rk["ip"] = [[100, 64, 0, 0], [0, 0, 64, 100], [192, 0, 2, 0]]
    sc.update("allowed_ranges_v4", rk, np.ones(3, np.uint8))

    sub = r.integers(0, n_subs + 8, n)  # >= n_subs: unknown MAC
    kind = r.integers(0, 12, n)
    macs = S.sub_mac_key(sub)
    src = S.sub_ip(sub)
    spoof = kind == 1
    src = np.where(spoof, src ^ np.uint32(0x00010000), src)
    src = np.where(kind == 2, np.uint32(0xC0000205), src)  # 192.0.2.5: only in the third range
    lens = np.full(n, 64, np.uint32)
    hdr = S.ipv4_headers(macs, np.uint64(GW_MAC), src, np.uint32(0x08080808), 6, 40000, 443, lens)
    is6 = (kind == 3) | (kind == 4)
    src6 = np.zeros((n, 16), np.uint8)
    src6[:, 0], src6[:, 1] = 0x20, 0x01
    src6[:, 15] = np.where(kind == 3, sub & 0xFF, (sub + 1) & 0xFF)
    h6 = ipv6_headers(macs, src6, n)
    hdr[is6] = h6[is6]
    arp = kind == 5
    hdr[arp, 12], hdr[arp, 13] = 0x08, 0x06
    vlan = kind == 6
    hdr[vlan, 12], hdr[vlan, 13] = 0x81, 0x00
    short = kind == 7
    lens[short] = r.integers(0, 40, short.sum())
    short6 = is6 & (r.integers(0, 6, n) == 0)
    lens[short6] = r.integers(34, 56, short6.sum())
    arena = fixed(hdr)
    sc.run("antispoof_ingress", arena, lens, 1_000_000_123)
    for dm, lg, now in ((2, 0, 2_000_000_000), (0, 1, 3_000_000_000), (3, 1, 4_000_000_000), (2, 1, 5_000_000_000)):
        cfg["default_mode"], cfg["log_violations"] = dm, lg
        sc.update1("antispoof_config", np.uint32(0), cfg)
        sc.run("antispoof_ingress", arena, lens, now)
    # control-plane round trips
    sc.lookup("subscriber_bindings", keys[3]).delete("subscriber_bindings", keys[3]).lookup("subscriber_bindings", keys[3])
    sc.delete("subscriber_bindings", keys[3])
    sc.update1("subscriber_bindings", keys[4], v[4], 1)  # NOEXIST on existing key
    sc.update1("subscriber_bindings", np.uint64(0x0200AABBCCDD), v[4], 2)  # EXIST on missing key
    sc.delete("allowed_ranges_v4", rk[1])
    sc.run("antispoof_ingress", arena, lens, 6_000_000_000)
    return sc


# ---------------------------------------------------------------------------
# qos_egress_prog / qos_ingress_prog
# ---------------------------------------------------------------------------
def qos_script(seed=0x905, n_subs=40, n=4000) -> Script:
    r = rng(seed)
    sc = Script("qos")
    keys = S.ip_bytes(S.sub_ip(np.arange(n_subs)))
    for m, shift in (("qos_ingress", 0), ("qos_egress", 3)):
        v = np.zeros(n_subs, L.token_bucket)
        rates = np.array([0, 8_000, 800_000, 80_000_000, 10_000_000_000, 64, 7, 1 << 40], dtype=np.uint64)
        bursts = np.array([1000, 1500, 3000, 65536, 10 << 20, 64, 200, 4000], dtype=np.uint32)
        i = (np.arange(n_subs) + shift) % 8
        v["rate_bps"], v["burst_bytes"] = rates[i], bursts[i]
        v["tokens"] = np.where(np.arange(n_subs) % 4 == 1, 0, bursts[i])
        v["tokens"][5] = 1 << 33  # above burst: clamped on first use
        v["priority"] = np.arange(n_subs) % 8
        v["last_update"] = np.where(np.arange(n_subs) % 7 == 2, 900_000_000, 0)
        sc.update(m, keys, v)
    sub = r.integers(0, n_subs + 4, n)
    sub = np.where(r.integers(0, 3, n) == 0, sub % 5, sub)  # a few hot subscribers
    lens = r.choice(np.array([64, 64, 64, 128, 594, 1518, 40, 34], dtype=np.uint32), n)
    hdr_in = S.ipv4_headers(S.sub_mac_key(sub), np.uint64(GW_MAC), S.sub_ip(sub), np.uint32(0x08080404), 17, 5000, 53, lens)
    hdr_eg = S.ipv4_headers(np.uint64(GW_MAC), S.sub_mac_key(sub), np.uint32(0x08080404), S.sub_ip(sub), 17, 53, 5000, lens)
    odd = r.integers(0, 40, n)
    for h in (hdr_in, hdr_eg):
        h[odd == 0, 12:14] = [0x86, 0xDD]
    lens2 = lens.copy()
    lens2[odd == 1] = r.integers(0, 34, (odd == 1).sum())
    ai, ae = fixed(hdr_in), fixed(hdr_eg)
    t = 1_000_000_000
    for step, dt in enumerate((0, 1_000_000, 250_000_000, 4_000_000_000, -3_000_000_000, 10**12, 1)):
        t += dt
        sc.run("qos_ingress_prog", ai, lens2, t)
        pr = np.full(n, 0xAAAA0000 + step, np.uint32)
        sc.run("qos_egress_prog", ae, lens2, t + 17, priority=pr)
    hdrs, l3 = hdr_in[:600], S.imix_lengths(600, seed)
    arena, off16 = S.pack_arena(hdrs, l3)
    sc.run("qos_ingress_prog", arena, l3, t + 10**9, off16=off16)
    sc.delete("qos_ingress", keys[0]).delete("qos_ingress", keys[0])
    sc.run("qos_ingress_prog", ai, lens2, t + 2 * 10**9)
    return sc


# ---------------------------------------------------------------------------
# nat44_egress / nat44_ingress / nat44_hairpin_xdp
# ---------------------------------------------------------------------------
def nat_maps(sc: Script, n_subs: int, pps: int, flags: int):
    keys, v, pubs = S.nat_blocks(n_subs, ports_per_sub=pps, port_lo=1024, port_hi=1024 + 8 * pps - 1)
    sc.update("subscriber_nat", keys, v)
    sc.update1("nat_config_map", np.uint32(0), S.nat_config(flags, pps))
    sc.update("hairpin_ips", S.ip_bytes(pubs), np.ones(len(pubs), np.uint8))
    alg = np.zeros(3, L.alg_config)
    alg["port"], alg["protocol"], alg["alg_type"] = [21, 5060, 5060], [6, 17, 6], [1, 2, 2]
    ak = ((alg["port"].astype(np.uint32) << 16) | alg["protocol"]).astype("<u4")
    sc.update("alg_ports", ak, alg)
    pool = np.zeros(len(pubs), L.nat_pool_entry)
    pool["public_ip"] = S.ip_bytes(pubs)
    pool["ports_per_sub"] = pps
    sc.update("nat_pool", np.arange(len(pubs), dtype="<u4"), pool)
    return pubs


def nat_frames(r, n_subs, n, pubs, n_flow_ports=6):
    sub = r.integers(0, n_subs, n)
    kind = r.integers(0, 40, n)
    proto = r.choice(np.array([6, 6, 6, 17, 17, 1], dtype=np.uint32), n)
    sport = (20000 + r.integers(0, n_flow_ports, n)).astype(np.uint32)
    dport = r.choice(np.array([443, 80, 53, 21, 5060, 8080], dtype=np.uint32), n)
    dst = (np.uint32(0x08080000) + r.integers(0, 3, n).astype(np.uint32)).astype(np.uint32)
    dst = np.where(kind == 1, np.uint32(pubs[0]), dst)  # hairpin target
    src = S.sub_ip(sub)
    src = np.where(kind == 2, np.uint32(0x08080101), src)  # public source: not NATed
    src = np.where(kind == 3, np.uint32(0x0A010101), src)  # private, no allocation
    proto = np.where(kind == 4, 47, proto)  # GRE: unsupported protocol
    lens = np.full(n, 64, np.uint32)
    ck = (r.integers(1, 65536, n)).astype(np.uint32)
    ck = np.where((proto == 17) & (kind % 5 == 0), 0, ck)  # UDP without checksum
    ck = np.where((proto == 17) & (kind == 6), 0xFFFF, ck)
    flags = r.choice(np.array([0x02, 0x10, 0x18, 0x11, 0x04], dtype=np.uint8), n)
    hdr = S.ipv4_headers(S.sub_mac_key(sub), np.uint64(GW_MAC), src, dst, proto, sport, dport, lens, l4_check=ck,
                         tcp_flags=flags)
    opt = kind == 7  # IPv4 options: ihl=6, L4 at 38
    if opt.any():
        h6 = S.ipv4_headers(S.sub_mac_key(sub), np.uint64(GW_MAC), src, dst, np.where(proto == 6, 17, proto), sport, dport,
                            lens, l4_check=ck, ihl=6)
        hdr[opt] = h6[opt]
    lens[kind == 8] = 50   # TCP header cut short
    lens[kind == 9] = 38   # UDP/ICMP header cut short
    lens[kind == 10] = 20  # IP header cut short
    hdr[kind == 11, 12:14] = [0x86, 0xDD]
    return hdr, lens


def nat_script(seed=0x44, flags=0x0F, n_subs=20, pps=8, n=2500, name="nat") -> Script:
    r = rng(seed)
    sc = Script(name)
    pubs = nat_maps(sc, n_subs, pps, flags)
    t = 5_000_000_000
    for b in range(3):
        hdr, lens = nat_frames(r, n_subs, n, pubs, n_flow_ports=3 + 3 * b)
        sc.run("nat44_egress", fixed(hdr), lens, t)
        sc.run("nat44_hairpin_xdp", fixed(hdr), lens, t)
        t += 1_000_000_000
    # IMIX layout through the offset table
    hdr, lens = nat_frames(r, n_subs, 800, pubs, n_flow_ports=4)
    l2 = np.where(lens == 64, S.imix_lengths(800, seed), lens)
    arena, off16 = S.pack_arena(hdr, l2)
    sc.run("nat44_egress", arena, l2, t, off16=off16)

    # ---- return traffic: replies to what the first egress batch sent ----
    first_run = next(i for i, st in enumerate(sc.steps) if st[0] == "run")

    def replies(res, variant):
        fr = res[f"s{first_run:03d}_frames"].reshape(-1, 64).copy()
        vd = res[f"s{first_run:03d}_verdict"]
        ln = res[f"s{first_run:03d}_len"]
        keep = (vd == 0) & (ln == 64) & (fr[:, 12] == 0x08) & (fr[:, 13] == 0x00) & ((fr[:, 14] & 0x0F) == 5)
        fr = fr[keep][:1500]
        out = fr.copy()
        out[:, 26:30], out[:, 30:34] = fr[:, 30:34], fr[:, 26:30]
        tcpudp = (fr[:, 23] == 6) | (fr[:, 23] == 17)
        out[tcpudp, 34:36], out[tcpudp, 36:38] = fr[tcpudp, 36:38], fr[tcpudp, 34:36]
        rr = rng(seed + variant)
        fl = rr.choice(np.array([0x12, 0x10, 0x10, 0x11, 0x04, 0x02], dtype=np.uint8), len(out))
        tcp = fr[:, 23] == 6
        out[tcp, 47] = fl[tcp]
        stray = rr.integers(0, 25, len(out)) == 0
        out[stray, 36] ^= 0x40  # unknown destination port: no reverse entry
        return {"arena": out.reshape(-1), "lens": np.full(len(out), 64, np.uint32),
                "now_ns": t + variant * 10**9, "stride": 64}

    sc.run_from("nat44_ingress", lambda res: replies(res, 1))
    sc.run_from("nat44_ingress", lambda res: replies(res, 2))

    sc.run_from("nat44_ingress", lambda res: replies(res, 3))
    return sc


def nat_stale_script(seed=0x57A) -> Script:
    """sessions removed underneath nat_reverse: 'sessions_expired' path of nat44_ingress (:871-876)."""
    r = rng(seed)
    sc = Script("nat_stale")
    n_subs = 6
    pubs = nat_maps(sc, n_subs, 64, 0x0F)
    sub = np.repeat(np.arange(n_subs), 8)
    sport = (30000 + np.tile(np.arange(8), n_subs)).astype(np.uint32)
    lens = np.full(len(sub), 64, np.uint32)
    hdr = S.ipv4_headers(S.sub_mac_key(sub), np.uint64(GW_MAC), S.sub_ip(sub), np.uint32(0x08080808), 6, sport, 443, lens,
                         l4_check=0x1234)
    sc.run("nat44_egress", fixed(hdr), lens, 10**9)
    # the control plane deletes half of the sessions (keys are known: 5-tuples of the frames)
    k = np.zeros(len(sub), L.nat_key)
    k["src_ip"] = S.ip_bytes(S.sub_ip(sub))
    k["dst_ip"] = S.ip_bytes(np.full(len(sub), 0x08080808, np.uint32))
    k["src_port"] = S.port_bytes(sport)
    k["dst_port"] = S.port_bytes(np.full(len(sub), 443))
    k["protocol"] = 6
    for i in range(0, len(sub), 2):
        sc.delete("nat_sessions", k[i])
    sc.delete("nat_sessions", k[0])  # second delete: -ENOENT

    def replies(res):
        run_tag = [kk for kk in res if kk.endswith("_frames")][0]
        fr = res[run_tag].reshape(-1, 64)
        out = fr.copy()
        out[:, 26:30], out[:, 30:34] = fr[:, 30:34], fr[:, 26:30]
        out[:, 34:36], out[:, 36:38] = fr[:, 36:38], fr[:, 34:36]
        out = np.concatenate([out, out, out[::3]], axis=0)  # each stale tuple shows up 2-3 times in one batch
        return {"arena": out.reshape(-1), "lens": np.full(len(out), 64, np.uint32), "now_ns": 2 * 10**9, "stride": 64}

    sc.run_from("nat44_ingress", replies)
    sc.run_from("nat44_ingress", replies)
    return sc


def nat_exhaust_script(flags=0x0F, name="nat_exhaust") -> Script:
    """Port exhaustion (TC_ACT_SHOT, :698-705).  allocate_port_from_block() probes eim_table with the
    HOST-order candidate as internal_port (:450-455), so a block only exhausts when the subscriber
    already has mappings whose network-order source port, loaded little-endian, equals every
    candidate — e.g. source ports 4 (bytes 00 04 -> 0x0400 = 1024) and 260 (01 04 -> 1025)
    for the block [1024, 1025] — or when the parity filter rejects every candidate."""
    sc = Script(name)
    n_subs = 6
    keys, v, pubs = S.nat_blocks(n_subs, ports_per_sub=2, port_lo=1024, port_hi=1024 + 2 * 8 - 1)
    v["block"]["port_start"][3:] = 1024 + 16 + np.arange(3)  # one-port blocks: parity can reject everything
    v["block"]["port_end"][3:] = v["block"]["port_start"][3:]
    v["block"]["next_port"][3:] = v["block"]["port_start"][3:]
    sc.update("subscriber_nat", keys, v)
    sc.update1("nat_config_map", np.uint32(0), S.nat_config(flags, 2))
    sub, sport, proto = [], [], []
    for s in range(n_subs):
        lo = int(v["block"]["port_start"][s])
        hi = int(v["block"]["port_end"][s])
        cand = [((p & 0xFF) << 8) | (p >> 8) for p in range(lo, hi + 1)]  # ports whose LE load is a candidate
        seq = cand + [5000, 5001, 5000, cand[0], 6001, 6002, 7003]
        for k, sp in enumerate(seq):
            sub.append(s)
            sport.append(sp)
            proto.append(6 if k % 4 != 3 else 17)
    sub, sport, proto = np.array(sub), np.array(sport, np.uint32), np.array(proto, np.uint32)
    order = np.argsort(np.arange(len(sub)) % 11, kind="stable")  # interleave subscribers
    sub, sport, proto = sub[order], sport[order], proto[order]
    lens = np.full(len(sub), 64, np.uint32)
    hdr = S.ipv4_headers(S.sub_mac_key(sub), np.uint64(GW_MAC), S.sub_ip(sub), np.uint32(0x08080808), proto, sport, 443, lens,
                         l4_check=0x4321)
    a = fixed(hdr)
    sc.run("nat44_egress", a, lens, 10**9)
    sc.run("nat44_egress", a, lens, 2 * 10**9)
    hdr2 = S.ipv4_headers(S.sub_mac_key(sub), np.uint64(GW_MAC), S.sub_ip(sub), np.uint32(0x08080809), proto, sport + 1, 80, lens,
                          l4_check=0x1111)
    sc.run("nat44_egress", fixed(hdr2), lens, 3 * 10**9)
    return sc


def nat_wrap_script() -> Script:
    """Long-lived sessions whose 64-bit packet / byte counters are about to cross 2^32 (and 2^33):
    struct nat_session counters are plain __u64 in the reference (bpf/nat44.c:131-136); a dataplane
    that keeps the low words apart from the high words has to carry exactly."""
    sc = Script("nat_wrap")
    n_subs, per = 4, 3
    pubs = nat_maps(sc, n_subs, 64, 0x0F)
    sub = np.repeat(np.arange(n_subs), per)
    nf = len(sub)
    sport = (40000 + np.arange(nf)).astype(np.uint32)
    nport = (1024 + 64 * sub + np.tile(np.arange(per), n_subs)).astype(np.uint32)
    proto = np.where(np.arange(nf) % 3 == 2, 17, 6).astype(np.uint32)
    dst = np.full(nf, 0x08080808, np.uint32)
    k = np.zeros(nf, L.nat_key)
    k["src_ip"], k["dst_ip"] = S.ip_bytes(S.sub_ip(sub)), S.ip_bytes(dst)
    k["src_port"], k["dst_port"] = S.port_bytes(sport), S.port_bytes(np.full(nf, 443))
    k["protocol"] = proto
    v = np.zeros(nf, L.nat_session)
    v["nat_ip"], v["nat_port"] = S.ip_bytes(np.full(nf, pubs[0], np.uint32)), S.port_bytes(nport)
    v["orig_ip"], v["orig_port"] = k["src_ip"], k["src_port"]
    v["dest_ip"], v["dest_port"] = k["dst_ip"], k["dst_port"]
    v["protocol"], v["state"] = proto, 1
    v["created"] = v["last_seen"] = 10**9
    M = 1 << 32
    # per flow: (packets, bytes) start values for both directions; 64-byte frames, 20 / 9 hits per flow
    starts = [(M - 1, M - 65), (M - 1, M - 64), (M - 4, (1 << 33) - 300), (2 * M - 20, M - 1), (M - 1, 0), (M - 1, M - 1),
              (M - 19, 3 * M - 64 * 20), (5, 7 * M - 1), (M - 2, M - 129), (0, 0), ((1 << 40) - 1, (1 << 48) - 1), (M - 20, M - 1280)]
    for d in ("out", "in"):
        v[f"packets_{d}"] = [s[0] for s in starts]
        v[f"bytes_{d}"] = [s[1] for s in starts]
    sc.update("nat_sessions", k, v)
    rk = np.zeros(nf, L.nat_key)
    rk["src_ip"], rk["dst_ip"] = k["dst_ip"], v["nat_ip"]
    rk["src_port"], rk["dst_port"] = k["dst_port"], v["nat_port"]
    rk["protocol"] = proto
    sc.update("nat_reverse", rk, k)
    lens1 = np.full(nf, 64, np.uint32)
    eg1 = S.ipv4_headers(S.sub_mac_key(sub), np.uint64(GW_MAC), S.sub_ip(sub), dst, proto, sport, 443, lens1, l4_check=0x2222,
                         tcp_flags=0x10)
    in1 = S.ipv4_headers(np.uint64(GW_MAC), S.sub_mac_key(sub), dst, np.full(nf, pubs[0], np.uint32), proto, 443, nport, lens1,
                         l4_check=0x3333, tcp_flags=0x10)
    sc.run("nat44_egress", fixed(eg1), lens1, 2 * 10**9)  # one frame per flow: the single-step carries
    sc.run("nat44_ingress", fixed(in1), lens1, 2 * 10**9)
    rep = np.tile(np.arange(nf), 19)
    lens = np.full(len(rep), 64, np.uint32)
    sc.run("nat44_egress", fixed(eg1[rep]), lens, 3 * 10**9)
    sc.run("nat44_ingress", fixed(in1[rep[: nf * 8]]), lens[: nf * 8], 3 * 10**9)
    for i in range(nf):
        sc.lookup("nat_sessions", k[i])
    return sc


# ---------------------------------------------------------------------------
# dhcp_fastpath_prog
# ---------------------------------------------------------------------------
def dhcp_frame(mac_key: int, *, msg_type=1, frame_len=362, opts_prefix=b"", vlan=(), giaddr=0, ciaddr=0, flags=0,
               op=1, magic=b"\x63\x82\x53\x63", dport=67, proto=17, ihl=5, extra_opts=b"", xid=0x1234ABCD,
               raw_opts=None) -> np.ndarray:
    """One DHCP request as the relay/client would send it.  `vlan` = tuple of (tpid, vid)."""
    eth = bytes.fromhex("ffffffffffff") + int(mac_key).to_bytes(6, "big")
    for tpid, vid in vlan:
        eth += tpid.to_bytes(2, "big") + (vid & 0x0FFF | 0x2000).to_bytes(2, "big")
    eth += b"\x08\x00"
    bootp = bytes([op, 1, 6, 1]) + xid.to_bytes(4, "big") + b"\x00\x05" + flags.to_bytes(2, "big")
    bootp += ciaddr.to_bytes(4, "big") + b"\x00" * 8 + giaddr.to_bytes(4, "big")
    bootp += int(mac_key).to_bytes(6, "big") + b"\x00" * 10
    bootp += b"S" * 64 + b"F" * 128 + magic
    if raw_opts is None:
        o = opts_prefix + bytes([53, 1, msg_type]) + extra_opts + bytes([55, 4, 1, 3, 15, 6, 255])
    else:
        o = raw_opts
    l3len = frame_len - len(eth)
    ip_hl = ihl * 4
    udp_len = l3len - ip_hl
    ip = bytes([0x40 | ihl, 0]) + max(l3len, 0).to_bytes(2, "big") + b"\x00\x01\x00\x00\x80" + bytes([proto]) + b"\x00\x00"
    ip += (b"\x00" * 4 if not giaddr else giaddr.to_bytes(4, "big")) + b"\xff\xff\xff\xff" + b"\x01" * (ip_hl - 20)
    udp = (68).to_bytes(2, "big") + dport.to_bytes(2, "big") + max(udp_len, 0).to_bytes(2, "big") + b"\xbe\xef"
    f = eth + ip + udp + bootp + o
    f = f[:frame_len] + b"\x00" * max(0, frame_len - len(f))
    return np.frombuffer(f, dtype=np.uint8).copy()


def dhcp_script(seed=0xD4C) -> Script:
    sc = Script("dhcp")
    n_subs = 64
    macs = S.sub_mac_key(np.arange(n_subs))
    pa = np.zeros(n_subs, L.pool_assignment)
    pa["pool_id"] = 1 + (np.arange(n_subs) % 5)  # pool 5 does not exist
    pa["allocated_ip"] = S.ip_bytes(np.uint32(0x0A000000) + 256 * (np.arange(n_subs) % 5) + 10 + np.arange(n_subs))
    pa["lease_expiry"] = np.where(np.arange(n_subs) % 11 == 10, 3, 1 << 62)
    pa["client_class"] = 1
    sc.update("subscriber_pools", macs, pa)
    pools = np.zeros(4, L.ip_pool)
    pools["network"] = S.ip_bytes(np.uint32(0x0A000000) + 256 * np.arange(4))
    pools["prefix_len"] = [24, 18, 0, 32]
    pools["gateway"] = S.ip_bytes(np.uint32(0x0A000001) + 256 * np.arange(4))
    pools["dns_primary"] = S.ip_bytes(np.array([0x08080808, 0x01010101, 0, 0x09090909], np.uint32))
    pools["dns_secondary"] = S.ip_bytes(np.array([0x08080404, 0, 0x08080404, 0x95707070], np.uint32))
    pools["lease_time"] = [3600, 86400, 0xFFFFFFFF, 7]
    sc.update("ip_pools", np.arange(1, 5, dtype="<u4"), pools)
    cfg = np.zeros(1, L.dhcp_server_config)
    cfg["server_mac"] = [[0x02, 0xAA, 0xBB, 0xCC, 0xDD, 0x01]]
    cfg["server_ip"] = [[10, 255, 0, 1]]
    cfg["interface_index"] = 3
    sc.update1("server_config", np.uint32(0), cfg)
    vk = np.zeros(3, L.vlan_key)
    vk["s_tag"], vk["c_tag"] = [100, 200, 300], [0, 7, 4095]
    vpa = pa[:3].copy()
    vpa["allocated_ip"] = S.ip_bytes(np.array([0x0A0A0A01, 0x0A0A0A02, 0x0A0A0A03], np.uint32))
    sc.update("vlan_subscriber_pools", vk, vpa)
    cids = [b"eth 0/1/2:100", b"X" * 32, b"c"]
    ck = np.zeros(len(cids), L.circuit_id_key)
    for i, c in enumerate(cids):
        ck["data"][i, :len(c)] = np.frombuffer(c, np.uint8)
    cpa = pa[:3].copy()
    cpa["allocated_ip"] = S.ip_bytes(np.array([0x0A141401, 0x0A141402, 0x0A141403], np.uint32))
    sc.update("circuit_id_subscribers", ck, cpa)
    sc.update1("circuit_id_map", np.uint64(0x1122334455667788), np.uint64(int(macs[2])))

    def opt82(cid: bytes, rid=b"rid"):
        sub = bytes([1, len(cid)]) + cid + bytes([2, len(rid)]) + rid
        return bytes([82, len(sub)]) + sub

    unknown = 0x02DEAD000001
    F = []
    for i in range(n_subs):
        F.append(dhcp_frame(macs[i], msg_type=1 if i % 5 == 0 else 3))
    F += [
        dhcp_frame(unknown), dhcp_frame(unknown, msg_type=3),
        dhcp_frame(macs[1], frame_len=342), dhcp_frame(macs[1], frame_len=345), dhcp_frame(macs[1], frame_len=346),
        dhcp_frame(macs[1], frame_len=350), dhcp_frame(macs[1], frame_len=1500), dhcp_frame(macs[1], frame_len=300),
        dhcp_frame(macs[1], frame_len=290), dhcp_frame(macs[1], frame_len=100), dhcp_frame(macs[1], frame_len=42),
        dhcp_frame(macs[1], frame_len=13), dhcp_frame(macs[1], frame_len=281), dhcp_frame(macs[1], frame_len=282),
        dhcp_frame(macs[1], frame_len=293), dhcp_frame(macs[1], frame_len=294),
    ]
    for pre in (b"\x00", b"\x00\x00", b"\x0c\x01a", b"\x0c\x02ab", b"\x0c\x03abc", b"\x0c\x04abcd", b"\x0c\x05abcde",
                b"\x3d\x07\x01" + b"\x02" * 6):
        F.append(dhcp_frame(macs[2], opts_prefix=pre))
    for mt in (2, 4, 5, 7, 8, 0):
        F.append(dhcp_frame(macs[2], msg_type=mt))
    F += [
        dhcp_frame(macs[2], op=2), dhcp_frame(macs[2], magic=b"\x63\x82\x53\x64"), dhcp_frame(macs[2], dport=68),
        dhcp_frame(macs[2], proto=6), dhcp_frame(macs[2], ihl=6, frame_len=380), dhcp_frame(macs[2], ihl=6, frame_len=346),
        dhcp_frame(macs[2], giaddr=0x0A630001), dhcp_frame(macs[2], giaddr=0x0A630001, flags=0x8000),
        dhcp_frame(macs[2], ciaddr=0x0A00000B), dhcp_frame(macs[2], ciaddr=0x0A00000B, flags=0x8000),
        dhcp_frame(macs[2], flags=0x8000), dhcp_frame(0, ciaddr=0), dhcp_frame(macs[4]), dhcp_frame(macs[10]),
        dhcp_frame(macs[3]), dhcp_frame(macs[2]),
    ]
    for vl in (((0x8100, 100),), ((0x88A8, 100),), ((0x8100, 101),), ((0x88A8, 200), (0x8100, 7)), ((0x88A8, 300), (0x8100, 4095)),
               ((0x88A8, 300), (0x8100, 1)), ((0x8100, 200), (0x8100, 7)), ((0x88A8, 200), (0x88A8, 7))):
        F.append(dhcp_frame(macs[6], vlan=vl, frame_len=380))
        F.append(dhcp_frame(unknown, vlan=vl, frame_len=346 + 4 * len(vl)))
        F.append(dhcp_frame(unknown, vlan=vl, frame_len=345 + 4 * len(vl)))
    F.append(dhcp_frame(macs[6], vlan=((0x8100, 100),), frame_len=17))
    F.append(dhcp_frame(macs[6], vlan=((0x88A8, 200), (0x8100, 7)), frame_len=21))
    for cid in cids + [b"nomatch", b"", b"Y" * 33]:
        F.append(dhcp_frame(unknown, extra_opts=opt82(cid)))  # option 82 right after option 53 (position 3)
        F.append(dhcp_frame(macs[7], extra_opts=opt82(cid)))
        for padn in (9, 12, 16, 17):  # option 82 at positions 12..19 and beyond
            F.append(dhcp_frame(unknown, extra_opts=bytes([12, padn - 5]) + b"h" * (padn - 5) + opt82(cid)))
    F.append(dhcp_frame(unknown, raw_opts=bytes([53, 1, 1, 82, 3, 1, 1, 65, 255])))  # opt82 too short (<4)
    F.append(dhcp_frame(unknown, raw_opts=bytes([53, 1, 1, 82, 200, 1, 1, 99, 255])))  # opt82 len runs past the frame
    frames = [np.asarray(f) for f in F]
    lens = np.array([len(f) for f in frames], np.uint32)
    width = int(((lens.max() + 15) // 16) * 16)
    hdrs = np.zeros((len(frames), width), np.uint8)
    for i, f in enumerate(frames):
        hdrs[i, :len(f)] = f
    arena, off16 = S.pack_arena(hdrs, lens)
    sc.run("dhcp_fastpath_prog", arena, lens, 2_500_000_000, off16=off16)
    cfg["server_ip"] = 0
    sc.update1("server_config", np.uint32(0), cfg)
    sc.run("dhcp_fastpath_prog", arena, lens, 9_999_999_999, off16=off16)
    sc.delete("subscriber_pools", macs[1]).delete("vlan_subscriber_pools", vk[0]).delete("circuit_id_subscribers", ck[0])
    sc.delete("ip_pools", np.uint32(2))
    sc.run("dhcp_fastpath_prog", arena, lens, 10**12, off16=off16)
    return sc


# ---------------------------------------------------------------------------
# pipeline_up
# ---------------------------------------------------------------------------
def pipeline_script(seed=0x919E, n_subs=30, n=3000, flags=0x0F, prog="pipeline_up") -> Script:
    r = rng(seed)
    sc = Script("pipeline" if prog == "pipeline_up" else prog)
    keys, v = S.bindings(n_subs)
    v["mode"] = np.where(np.arange(n_subs) % 9 == 8, 3, 1)
    sc.update("subscriber_bindings", keys, v)
    cfg = np.zeros(1, L.antispoof_config)
    cfg["default_mode"], cfg["log_violations"] = 1, 1
    sc.update1("antispoof_config", np.uint32(0), cfg)
    pubs = nat_maps(sc, n_subs - 3, 16, flags)  # the last 3 subscribers have no NAT allocation
    qk, qv = S.qos_buckets(n_subs - 2)           # ... and the last 2 no bucket
    qv["burst_bytes"] = np.minimum(qv["burst_bytes"], 20000)
    qv["tokens"] = qv["burst_bytes"]
    sc.update("qos_ingress", qk, qv)
    sc.update("qos_egress", qk, qv)
    t = 10**9
    for b in range(4):
        hdr, lens = nat_frames(r, n_subs, n, pubs, n_flow_ports=2 + 2 * b)
        sub_unknown = r.integers(0, 30, n) == 0
        hdr[sub_unknown, 8] ^= 0x55  # unknown source MAC
        l2 = np.where(lens == 64, S.imix_lengths(n, seed + b), lens)
        arena, off16 = S.pack_arena(hdr, l2)
        sc.run(prog, arena, l2, t, off16=off16)
        t += 3_000_000
    hdr, lens = nat_frames(r, n_subs, n, pubs, n_flow_ports=8)
    sc.run(prog, fixed(hdr), lens, t + 10**9, stride=64)
    return sc


# ---------------------------------------------------------------------------
# IPv4 header lengths 0..15: L4 header anywhere from byte 14 (overlapping the IP header) to byte 74
# ---------------------------------------------------------------------------
def ihl_frames(r, n_subs, n, dst_base=0x08080000, width=128):
    """u8[n,width] frames whose IPv4 header claims every ihl 0..15 (bpf/nat44.c:606 multiplies whatever the nibble
    says), TCP / UDP / ICMP, long enough that the L4 header is present wherever it lands.  The L4 fields are
    written first and the fixed IPv4 fields over them, so for ihl < 5 — where the two overlap — the source
    address survives and the frame still reaches the NAT code."""
    sub = r.integers(0, n_subs, n)
    ihl = (np.arange(n) % 16).astype(np.int64)
    r.shuffle(ihl)
    proto = r.choice(np.array([6, 6, 17, 17, 1], dtype=np.uint32), n)
    sport = (20000 + r.integers(0, 5, n)).astype(np.uint32)
    dport = r.choice(np.array([443, 80, 53, 21, 5060], dtype=np.uint32), n)
    dst = (np.uint32(dst_base) + r.integers(0, 3, n).astype(np.uint32)).astype(np.uint32)
    base = S.ipv4_headers(S.sub_mac_key(sub), np.uint64(GW_MAC), S.sub_ip(sub), dst, np.uint32(47), 0, 0, np.full(n, width, np.uint32))
    f = np.zeros((n, width), np.uint8)
    f[:, 34:] = r.integers(0, 256, (n, width - 34), dtype=np.uint8)  # option bytes / payload: arbitrary
    l4 = 14 + 4 * ihl
    rows = np.arange(n)
    ck = r.integers(0, 65536, n).astype(np.uint32)
    ck = np.where((proto == 17) & (r.integers(0, 4, n) == 0), 0, ck)
    for k, (val, tcpudp_off, icmp_off) in enumerate(((sport, 0, 4), (dport, 2, None))):
        b = S.port_bytes(val)
        for j in range(2):
            m = proto != 1
            f[rows[m], l4[m] + tcpudp_off + j] = b[m, j]
            if icmp_off is not None:
                m = proto == 1
                f[rows[m], l4[m] + icmp_off + j] = b[m, j]
    ckb = S.port_bytes(ck)
    for j in range(2):
        for pr, off in ((6, 16), (17, 6), (1, 2)):
            m = proto == pr
            f[rows[m], l4[m] + off + j] = ckb[m, j]
    m = proto == 6
    f[rows[m], l4[m] + 13] = r.choice(np.array([0x02, 0x10, 0x18, 0x11, 0x04], dtype=np.uint8), int(m.sum()))
    f[:, 0:34] = base[:, 0:34]
    f[:, 14] = (0x40 | ihl).astype(np.uint8)
    f[:, 23] = proto
    f[:, 24:26] = 0
    f[:, 24:26] = S.ip_checksum(f[:, 14:34])
    return f, ihl, proto


def ipopts_script(seed=0x0B75, n_subs=12, n=1600) -> Script:
    """Frames with IPv4 options (and with header lengths below 5) through every TC program and every arena layout:
    the L4 header of a long-option frame lies past byte 63, i.e. outside the 64-byte header slot most frames need."""
    r = rng(seed)
    sc = Script("ipopts")
    keys, v = S.bindings(n_subs)
    sc.update("subscriber_bindings", keys, v)
    cfg = np.zeros(1, L.antispoof_config)
    cfg["default_mode"], cfg["log_violations"] = 1, 1
    sc.update1("antispoof_config", np.uint32(0), cfg)
    nat_maps(sc, n_subs, 256, 0x0F)
    qk, qv = S.qos_buckets(n_subs)
    qv["burst_bytes"] = np.minimum(qv["burst_bytes"], 30000)
    qv["tokens"] = qv["burst_bytes"]
    sc.update("qos_ingress", qk, qv)
    sc.update("qos_egress", qk, qv)
    t = 3 * 10**9
    f, ihl, proto = ihl_frames(r, n_subs, n)
    l4 = 14 + 4 * ihl
    lens = r.choice(np.array([128, 128, 594, 1518], dtype=np.uint32), n)
    cut = r.integers(0, 8, n)  # a slice of frames ends right at / just before the end of its L4 header
    need = l4 + np.where(proto == 6, 20, 8)
    lens = np.where(cut == 0, need, lens)
    lens = np.where(cut == 1, need - 1, lens).astype(np.uint32)
    # (a) packed arena, neighbours 16-byte adjacent; (b) 128-byte slots
    arena, off16 = S.pack_arena(f, lens)
    sc.run("nat44_egress", arena, lens, t, off16=off16)
    sc.run("nat44_egress", arena, lens, t + 10**6, off16=off16)
    l128 = np.minimum(lens, 128).astype(np.uint32)
    sc.run("nat44_egress", fixed(f, 128), l128, t + 2 * 10**6, stride=128)
    first_run = next(i for i, st in enumerate(sc.steps) if st[0] == "run")

    def replies(res):
        out_a = res[f"s{first_run:03d}_frames"]
        fr = np.stack([out_a[int(o) * 16: int(o) * 16 + 128] for o in off16])
        keep = (ihl >= 5) & (lens >= need) & (res[f"s{first_run:03d}_verdict"] == 0)
        fr, k_l4, k_pr = fr[keep], l4[keep], proto[keep]
        out = fr.copy()
        out[:, 26:30], out[:, 30:34] = fr[:, 30:34], fr[:, 26:30]
        rows = np.arange(len(out))
        for j in range(2):
            m = k_pr != 1
            out[rows[m], k_l4[m] + j], out[rows[m], k_l4[m] + 2 + j] = fr[rows[m], k_l4[m] + 2 + j], fr[rows[m], k_l4[m] + j]
        m = k_pr == 1
        out[rows[m], k_l4[m]] = 0  # echo reply
        ln = np.where(np.arange(len(out)) % 3 == 0, 594, 128).astype(np.uint32)
        a, o16 = S.pack_arena(out, ln)
        return {"arena": a, "lens": ln, "now_ns": t + 5 * 10**6, "off16": o16}

    sc.run_from("nat44_ingress", replies)
    sc.run_from("nat44_ingress", replies)
    f2, ihl2, proto2 = ihl_frames(r, n_subs, n, dst_base=0x08080400)
    lens2 = r.choice(np.array([110, 128, 594, 1518], dtype=np.uint32), n)
    spoof = r.integers(0, 20, n) == 0
    f2[spoof, 27] ^= 0x20
    arena2, off2 = S.pack_arena(f2, lens2)
    sc.run("pipeline_up", arena2, lens2, t + 10**7, off16=off2)
    sc.run("pipeline_up", arena2, lens2, t + 10**7 + 10**6, off16=off2)
    sc.run("pipeline_tc", arena2, lens2, t + 10**7 + 2 * 10**6, off16=off2)
    sc.run("antispoof_ingress", arena2, lens2, t + 2 * 10**7, off16=off2)
    sc.run("qos_ingress_prog", arena2, lens2, t + 3 * 10**7, off16=off2)
    sc.run("qos_egress_prog", arena2, lens2, t + 3 * 10**7, off16=off2, priority=np.zeros(n, np.uint32))
    sc.run("nat44_hairpin_xdp", arena2, lens2, t + 3 * 10**7, off16=off2)
    return sc



# ---------------------------------------------------------------------------
# a clock value per frame (the reference calls bpf_ktime_get_ns() per packet: bpf/nat44.c:669, bpf/qos_ratelimit.c:80)
# ---------------------------------------------------------------------------
def ticks_script(seed=0x71C5, n_subs=24, n=3000) -> Script:
    """Every program with per-frame timestamps: token buckets refill BETWEEN the frames of a batch (a frame that
    found the bucket empty passes a few microseconds later), sessions / mappings / log records carry their own
    frame's time, the lease check of the DHCP fast path too."""
    r = rng(seed)
    sc = Script("ticks")
    keys, v = S.bindings(n_subs)
    sc.update("subscriber_bindings", keys, v)
    cfg = np.zeros(1, L.antispoof_config)
    cfg["default_mode"], cfg["log_violations"] = 1, 1
    sc.update1("antispoof_config", np.uint32(0), cfg)
    pubs = nat_maps(sc, n_subs, 64, 0x0F)
    qk, qv = S.qos_buckets(n_subs)
    qv["burst_bytes"] = np.minimum(qv["burst_bytes"], 6000)  # small bursts: the refill between frames decides
    qv["tokens"] = qv["burst_bytes"] // 2
    qv["rate_bps"] = np.where(qv["rate_bps"] > 0, np.minimum(qv["rate_bps"], 1_000_000 * (1 + np.arange(n_subs) % 5)), 0)
    sc.update("qos_ingress", qk, qv)
    sc.update("qos_egress", qk, qv)
    t = 20 * 10**9
    first_nat = None
    for b in range(3):
        hdr, lens = nat_frames(r, n_subs, n, pubs, n_flow_ports=3 + 2 * b)
        spoof = r.integers(0, 25, n) == 0
        hdr[spoof, 27] ^= 0x08
        l2 = np.where(lens == 64, S.imix_lengths(n, seed + b), lens)
        arena, off16 = S.pack_arena(hdr, l2)
        now_v = (t + np.cumsum(r.integers(0, 40_000, n))).astype(np.uint64)  # 0..40 us between frames, repeats included
        prog = ("pipeline_up", "nat44_egress", "pipeline_tc")[b]
        if prog == "nat44_egress":
            first_nat = len(sc.steps)
        sc.run(prog, arena, l2, int(now_v[0]), off16=off16, now_v=now_v)
        sc.run("qos_ingress_prog", arena, l2, int(now_v[-1]), off16=off16, now_v=(now_v + np.uint64(10**6)))
        sc.run("qos_egress_prog", arena, l2, int(now_v[-1]), off16=off16, now_v=(now_v + np.uint64(2 * 10**6)),
               priority=np.zeros(n, np.uint32))
        sc.run("antispoof_ingress", arena, l2, int(now_v[-1]), off16=off16, now_v=(now_v + np.uint64(3 * 10**6)))
        t = int(now_v[-1]) + 5 * 10**6

    def replies(res):
        out_a = res[f"s{first_nat:03d}_frames"]
        st = sc.steps[first_nat]
        off16, lens = st[5], st[3]
        fr = np.stack([out_a[int(o) * 16: int(o) * 16 + 64] for o in off16])
        keep = (res[f"s{first_nat:03d}_verdict"] == 0) & (lens >= 64) & (fr[:, 12] == 0x08) & ((fr[:, 14] & 0x0F) == 5)
        fr = fr[keep][:1200]
        out = fr.copy()
        out[:, 26:30], out[:, 30:34] = fr[:, 30:34], fr[:, 26:30]
        tcpudp = (fr[:, 23] == 6) | (fr[:, 23] == 17)
        out[tcpudp, 34:36], out[tcpudp, 36:38] = fr[tcpudp, 36:38], fr[tcpudp, 34:36]
        out[fr[:, 23] == 6, 47] = 0x10
        rr = rng(seed + 99)
        nv = (t + np.cumsum(rr.integers(0, 9_000, len(out)))).astype(np.uint64)
        return {"arena": out.reshape(-1), "lens": np.full(len(out), 64, np.uint32), "now_ns": int(nv[0]), "stride": 64, "now_v": nv}

    sc.run_from("nat44_ingress", replies)
    # DHCP: leases that expire WHILE the batch runs
    macs = S.sub_mac_key(np.arange(16))
    pa = np.zeros(16, L.pool_assignment)
    pa["pool_id"], pa["allocated_ip"], pa["client_class"] = 1, S.ip_bytes(np.uint32(0x0A000010) + np.arange(16)), 1
    pa["lease_expiry"] = 100 + np.arange(16) % 4  # seconds
    sc.update("subscriber_pools", macs, pa)
    pools = np.zeros(1, L.ip_pool)
    pools["network"], pools["prefix_len"], pools["gateway"] = S.ip_bytes(np.uint32(0x0A000000)), 24, S.ip_bytes(np.uint32(0x0A000001))
    pools["dns_primary"], pools["lease_time"] = S.ip_bytes(np.uint32(0x08080808)), 600
    sc.update("ip_pools", np.ones(1, "<u4"), pools)
    dcfg = np.zeros(1, L.dhcp_server_config)
    dcfg["server_mac"], dcfg["server_ip"] = [[2, 0xAA, 0xBB, 0xCC, 0xDD, 1]], [[10, 0, 0, 1]]
    sc.update1("server_config", np.uint32(0), dcfg)
    F = [dhcp_frame(macs[i % 16], msg_type=3) for i in range(400)]
    lens = np.array([len(f) for f in F], np.uint32)
    width = int(((lens.max() + 15) // 16) * 16)
    hd = np.zeros((len(F), width), np.uint8)
    for i, f in enumerate(F):
        hd[i, :len(f)] = f
    arena, off16 = S.pack_arena(hd, lens)
    nv = (99 * 10**9 + np.arange(400, dtype=np.uint64) * np.uint64(15 * 10**6))  # 99.0 s .. 105 s
    sc.run("dhcp_fastpath_prog", arena, lens, int(nv[0]), off16=off16, now_v=nv)
    return sc


ALL_SCRIPTS = {
    "antispoof": antispoof_script,
    "qos": qos_script,
    "nat": nat_script,
    "nat_parity": lambda: nat_script(seed=0x45, flags=0x2F, name="nat_parity"),
    "nat_noeim": lambda: nat_script(seed=0x46, flags=0x0E, name="nat_noeim"),
    "nat_sip": lambda: nat_script(seed=0x47, flags=0x1D, pps=32, name="nat_sip"),
    "nat_stale": nat_stale_script,
    "nat_wrap": nat_wrap_script,
    "nat_exhaust": nat_exhaust_script,
    "nat_exhaust_parity": lambda: nat_exhaust_script(flags=0x2F, name="nat_exhaust_parity"),
    "nat_exhaust_noeim": lambda: nat_exhaust_script(flags=0x2E, name="nat_exhaust_noeim"),
    "dhcp": dhcp_script,
    "pipeline": pipeline_script,
    "pipeline_noeim": lambda: pipeline_script(seed=0x91A0, flags=0x06),
    "ipopts": ipopts_script,
    "ticks": ticks_script,
    # the order the reference's TC hooks give: antispoof -> qos_ingress -> nat44_egress (a frame the bucket drops never reaches NAT)
    "pipeline_tc": lambda: pipeline_script(seed=0x91A7, prog="pipeline_tc"),
    "pipeline_tc_noeim": lambda: pipeline_script(seed=0x91A9, flags=0x06, prog="pipeline_tc"),
}

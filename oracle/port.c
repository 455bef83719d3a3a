/* TEST INFRASTRUCTURE — plain-C restatement of the reference's four dataplane
 * programs, written against frame byte offsets and the eBPF-map emulation in
 * oracle/runtime.c.  It exists so that a CPU checker is available wherever the
 * reference sources are not (oracle/_ref cannot be rebuilt without
 * /root/reference), and as an independent second reading of the algorithm.
 *
 * Pinned: tests/test_oracle_golden.py requires this library to reproduce,
 * bit for bit, the golden vectors in tests/golden/ that were produced by the
 * reference's own C run natively (tests/golden/make_golden.py).
 *
 * Every function cites the reference lines it restates (paths relative to
 * the reference tree).  All multi-byte frame fields are little-endian loads of
 * wire bytes, as the eBPF programs see them.
 */
#include <string.h>

#include "oracle_api.h"

typedef uint8_t u8;
typedef uint16_t u16;
typedef uint32_t u32;
typedef uint64_t u64;

/* ---- map table: names, types, key/value sizes, capacities of the reference ----
 * bpf/antispoof.c:71-119, bpf/qos_ratelimit.c:37-65, bpf/nat44.c:218-320, bpf/maps.h:99-234 */
enum {
    M_BINDINGS, M_AS_CFG, M_AS_STATS, M_SPOOF_EV, M_RANGES, M_QOS_EG, M_QOS_IN, M_QOS_STATS, M_SESSIONS, M_REVERSE,
    M_EIM, M_SUBNAT, M_POOL, M_HAIRPIN, M_NAT_CFG, M_NAT_STATS, M_NAT_LOG, M_ALG, M_PRIV, M_SUBPOOLS, M_VLANPOOLS,
    M_IPPOOLS, M_SRV_CFG, M_DHCP_STATS, M_CIDMAP, M_CIDSUBS, M_COUNT
};
static const ora_map_desc k_maps[M_COUNT] = {
    {"subscriber_bindings", ORA_MAP_HASH, 8, 24, 1000000, 0},
    {"antispoof_config", ORA_MAP_ARRAY, 4, 8, 1, 0},
    {"antispoof_stats", ORA_MAP_PERCPU_ARRAY, 4, 48, 1, 0},
    {"spoof_events", ORA_MAP_PERF_EVENT_ARRAY, 4, 4, 0, 0},
    {"allowed_ranges_v4", ORA_MAP_LPM_TRIE, 8, 1, 256, 0},
    {"qos_egress", ORA_MAP_HASH, 4, 32, 1000000, 0},
    {"qos_ingress", ORA_MAP_HASH, 4, 32, 1000000, 0},
    {"qos_stats_map", ORA_MAP_PERCPU_ARRAY, 4, 32, 1, 0},
    {"nat_sessions", ORA_MAP_LRU_HASH, 16, 80, 4000000, 0},
    {"nat_reverse", ORA_MAP_LRU_HASH, 16, 16, 4000000, 0},
    {"eim_table", ORA_MAP_LRU_HASH, 8, 32, 2000000, 0},
    {"subscriber_nat", ORA_MAP_HASH, 4, 64, 1000000, 0},
    {"nat_pool", ORA_MAP_ARRAY, 4, 16, 256, 0},
    {"hairpin_ips", ORA_MAP_HASH, 4, 1, 1000, 0},
    {"nat_config_map", ORA_MAP_ARRAY, 4, 16, 1, 0},
    {"nat_stats_map", ORA_MAP_PERCPU_ARRAY, 4, 104, 1, 0},
    {"nat_log_rb", ORA_MAP_RINGBUF, 0, 0, 1 << 20, 0},
    {"alg_ports", ORA_MAP_HASH, 4, 8, 64, 0},
    {"nat_private_ranges", ORA_MAP_LPM_TRIE, 8, 1, 64, 0},
    {"subscriber_pools", ORA_MAP_HASH, 8, 25, 1000000, 0},
    {"vlan_subscriber_pools", ORA_MAP_HASH, 4, 25, 100000, 0},
    {"ip_pools", ORA_MAP_HASH, 4, 28, 10000, 0},
    {"server_config", ORA_MAP_ARRAY, 4, 16, 1, 0},
    {"stats_map", ORA_MAP_ARRAY, 4, 80, 1, 0},
    {"circuit_id_map", ORA_MAP_HASH, 8, 8, 1000000, 0},
    {"circuit_id_subscribers", ORA_MAP_HASH, 32, 25, 1000000, 0},
};
static struct ora_map *mp[M_COUNT];

/* ---- little helpers ---- */
static inline u16 ld16(const u8 *f, u32 o) { return (u16)(f[o] | (f[o + 1] << 8)); }
static inline u32 ld32(const u8 *f, u32 o) { return (u32)f[o] | ((u32)f[o + 1] << 8) | ((u32)f[o + 2] << 16) | ((u32)f[o + 3] << 24); }
static inline u64 ld64(const u8 *f) { return (u64)ld32(f, 0) | ((u64)ld32(f, 4) << 32); }
static inline void st16(u8 *f, u32 o, u16 v) { f[o] = (u8)v; f[o + 1] = (u8)(v >> 8); }
static inline void st32(u8 *f, u32 o, u32 v) { st16(f, o, (u16)v); st16(f, o + 2, (u16)(v >> 16)); }
static inline void st64(u8 *f, u64 v) { st32(f, 0, (u32)v); st32(f, 4, (u32)(v >> 32)); }
static inline u16 swab16(u16 x) { return (u16)((x << 8) | (x >> 8)); }
static inline u32 swab32(u32 x) { return __builtin_bswap32(x); }
static inline u64 *stat(int m, int idx) {
    u32 zero = 0;
    u8 *s = ora_rt_lookup(mp[m], &zero);
    return (u64 *)s + idx;
}
static u64 mac48(const u8 *m) { /* bpf/antispoof.c:122-129, bpf/dhcp_fastpath.c:175-182 */
    u64 k = 0;
    for (int i = 0; i < 6; i++) k = (k << 8) | m[i];
    return k;
}

/* ======================= antispoof_ingress, bpf/antispoof.c:188-293 ======================= */
static void spoof_log(const u8 *f, u32 spoofed, u32 allowed, int v6) { /* log_violation(), :150-175 */
    u8 ev[56];
    memset(ev, 0, sizeof(ev));
    st64(ev, ora_rt_now());
    memcpy(ev + 8, f + 6, 6);
    ev[14] = v6 ? 6 : 4;
    if (!v6) {
        st32(ev, 16, spoofed);
        st32(ev, 20, allowed);
    }
    ora_rt_event_output(mp[M_SPOOF_EV], ev, sizeof(ev));
    ++*stat(M_AS_STATS, 2);
}

static int p_antispoof(ora_pkt *p) {
    const u8 *f = p->data;
    u32 len = p->len;
    if (len < 14) return 0;
    u64 key = mac48(f + 6);
    u32 zero = 0;
    const u8 *cfg = ora_rt_lookup(mp[M_AS_CFG], &zero);
    u8 defmode = cfg ? cfg[0] : 0, logv = cfg ? cfg[1] : 0;
    const u8 *b = ora_rt_lookup(mp[M_BINDINGS], &key); /* ipv4@0 ipv6@4 v4valid@20 v6valid@21 mode@22 */
    u8 mode = b ? b[22] : defmode;
    u64 *allowed_ctr = stat(M_AS_STATS, 0), *dropped = stat(M_AS_STATS, 1);
    if (mode == 0) {
        ++*allowed_ctr;
        return 0;
    }
    u16 et = ld16(f, 12);
    if (et == 0x0008) {
        if (len < 34) return 0;
        u32 src = ld32(f, 26);
        int ok = 0;
        if (b && b[20]) {
            if (mode == 1 || mode == 3) ok = src == ld32(b, 0);
        } else if (mode == 2) {
            u8 lk[8];
            st32(lk, 0, 32);
            st32(lk, 4, src);
            ok = ora_rt_lookup(mp[M_RANGES], lk) != NULL; /* ip_in_allowed_range(), :178-185 */
        }
        if (!ok) {
            if (logv) spoof_log(f, src, b ? ld32(b, 0) : 0, 0);
            if (mode == 3) {
                ++*allowed_ctr;
                return 0;
            }
            ++*dropped;
            ++*stat(M_AS_STATS, 3);
            return 2;
        }
        ++*allowed_ctr;
        return 0;
    }
    if (et == 0xDD86) {
        if (len < 54) return 0;
        int ok = 0;
        if (b && b[21])
            ok = memcmp(f + 22, b + 4, 16) == 0;
        else if (mode == 2)
            ok = 1;
        if (!ok && mode != 3) {
            if (logv) spoof_log(f, 0, 0, 1);
            ++*dropped;
            ++*stat(M_AS_STATS, 4);
            return 2;
        }
        ++*allowed_ctr;
        return 0;
    }
    ++*allowed_ctr;
    return 0;
}

/* ======================= QoS, bpf/qos_ratelimit.c ======================= */
/* token_bucket: tokens@0 last_update@8 rate_bps@16 burst_bytes@24 priority@28; token_bucket_check(), :70-104 */
static int bucket_take(u8 *tb, u32 pkt_len) {
    u64 rate = ld64(tb + 16);
    if (rate == 0) return 1;
    u64 now = ora_rt_now();
    u64 tokens = ld64(tb) + ((now - ld64(tb + 8)) * (rate / 8)) / 1000000000ULL;
    u32 burst = ld32(tb, 24);
    if (tokens > burst) tokens = burst;
    int ok = tokens >= pkt_len;
    if (ok) tokens -= pkt_len;
    st64(tb, tokens);
    st64(tb + 8, now);
    return ok;
}

static int qos_common(ora_pkt *p, int egress) { /* :126-172 and :178-222 */
    const u8 *f = p->data;
    if (p->len < 14 || ld16(f, 12) != 0x0008 || p->len < 34) return 0;
    u32 ip = ld32(f, egress ? 30 : 26);
    u8 *tb = ora_rt_lookup(mp[egress ? M_QOS_EG : M_QOS_IN], &ip);
    if (!tb) return 0;
    int ok = bucket_take(tb, p->len);
    *stat(M_QOS_STATS, ok ? 0 : 1) += 1;
    *stat(M_QOS_STATS, ok ? 2 : 3) += p->len;
    if (!ok) return 2;
    if (egress) p->priority = tb[28];
    return 0;
}
static int p_qos_egress(ora_pkt *p) { return qos_common(p, 1); }
static int p_qos_ingress(ora_pkt *p) { return qos_common(p, 0); }

/* ======================= NAT44, bpf/nat44.c ======================= */
enum { NS_SNAT, NS_DNAT, NS_HAIRPIN, NS_DROPPED, NS_PASSED, NS_CREATED, NS_EXPIRED, NS_EXHAUST, NS_EIM_HIT, NS_EIM_MISS, NS_ALG };

static int private_src(u32 ip) { /* is_private_ip(), :340-363 */
    u32 a = ip & 0xff, b = (ip >> 8) & 0xff;
    return a == 10 || (a == 172 && b >= 16 && b <= 31) || (a == 192 && b == 168) || (a == 100 && b >= 64 && b <= 127);
}
static u16 fold(u32 s) { /* csum_fold(), :378-382 */
    s = (s & 0xffff) + (s >> 16);
    s = (s & 0xffff) + (s >> 16);
    return (u16)~s;
}
static u16 csum32(u16 c, u32 o, u32 n) { /* update_csum(), :384-391 */
    u32 s = ~((u32)c) & 0xffff;
    s += ~o & 0xffff;
    s += ~(o >> 16) & 0xffff;
    s += n & 0xffff;
    s += n >> 16;
    return fold(s);
}
static u16 csum16(u16 c, u16 o, u16 n) { /* update_csum16(), :393-398 */
    u32 s = ~((u32)c) & 0xffff;
    s += ~(u32)o & 0xffff;
    s += n;
    return fold(s);
}
static void natlog(u32 type, u32 sub_id, u32 priv, u32 pub, u16 pport, u16 pubport, u32 dst, u16 dport, u8 proto, u8 fl) {
    u8 *e = ora_rt_ringbuf_reserve(mp[M_NAT_LOG], 40); /* log_nat_event(), :531-562 */
    if (!e) return;
    st64(e, ora_rt_now());
    st32(e, 8, type);
    st32(e, 12, sub_id);
    st32(e, 16, priv);
    st32(e, 20, pub);
    st16(e, 24, pport);
    st16(e, 26, pubport);
    st32(e, 28, dst);
    st16(e, 32, dport);
    e[34] = proto;
    e[35] = fl;
    ora_rt_ringbuf_submit(e);
}
static void eim_key(u8 *k, u32 ip, u16 port, u8 proto) {
    st32(k, 0, ip);
    st16(k, 4, port);
    k[6] = proto;
    k[7] = 0;
}
/* port_block inside subscriber_nat: public_ip@0 port_start@4 port_end@6 next_port@8 subscriber_id@24;
 * allocate_port_from_block(), :408-466 */
static u16 take_port(u8 *blk, int parity, u16 orig_port, u32 ip, u8 proto) {
    u32 lo = ld16(blk, 4), hi = ld16(blk, 6);
    for (int i = 0; i < 64; i++) {
        u32 next = ld32(blk, 8);
        u16 port = (u16)next;
        st32(blk, 8, next + 1);
        if (port > hi) port = (u16)lo;
        if (ld32(blk, 8) > hi) st32(blk, 8, lo);
        if (parity && ((port ^ orig_port) & 1)) continue;
        u8 k[8];
        eim_key(k, ip, port, proto);
        if (ora_rt_lookup(mp[M_EIM], k)) continue;
        return port;
    }
    return 0;
}

static int p_nat_egress(ora_pkt *p) { /* nat44_egress, :565-802 */
    u8 *f = p->data;
    u32 len = p->len, zero = 0;
    if (len < 14 || ld16(f, 12) != 0x0008 || len < 34) return 0;
    u32 sip = ld32(f, 26), dip = ld32(f, 30);
    if (!private_src(sip)) return 0;
    u8 *sub = ora_rt_lookup(mp[M_SUBNAT], &sip);
    if (!sub) {
        ++*stat(M_NAT_STATS, NS_PASSED);
        return 0;
    }
    const u8 *cfg = ora_rt_lookup(mp[M_NAT_CFG], &zero);
    u32 fl = cfg ? ld32(cfg, 0) : 0;
    u8 proto = f[23];
    u32 l4 = 14 + (f[14] & 15) * 4;
    u16 sp = 0, dp = 0;
    if (proto == 6 || proto == 17) {
        if (l4 + (proto == 6 ? 20 : 8) > len) return 0;
        sp = ld16(f, l4);
        dp = ld16(f, l4 + 2);
        if (fl & (proto == 6 ? 0x18 : 0x10)) { /* check_alg_trigger(), :372-375 */
            u32 ak = ((u32)swab16(dp) << 16) | proto;
            const u8 *alg = ora_rt_lookup(mp[M_ALG], &ak);
            if (alg) {
                ++*stat(M_NAT_STATS, NS_ALG);
                natlog(7, ld32(sub, 24), sip, 0, sp, 0, dip, dp, proto, alg[3]);
                return 0;
            }
        }
    } else if (proto == 1) {
        if (l4 + 8 > len) return 0;
        sp = ld16(f, l4 + 4);
    } else {
        return 0;
    }
    u8 hairpin = 0;
    if ((fl & 4) && ora_rt_lookup(mp[M_HAIRPIN], &dip)) {
        hairpin = 1;
        ++*stat(M_NAT_STATS, NS_HAIRPIN);
    }
    u8 key[16];
    memset(key, 0, 16);
    st32(key, 0, sip);
    st32(key, 4, dip);
    st16(key, 8, sp);
    st16(key, 10, dp);
    key[12] = proto;
    u64 now = ora_rt_now();
    u8 *ses = ora_rt_lookup(mp[M_SESSIONS], key);
    u32 nip;
    u16 nport;
    if (ses) { /* nat_session: nat_ip@0 nat_port@4 last_seen@24 packets_out@40 bytes_out@56 */
        nip = ld32(ses, 0);
        nport = ld16(ses, 4);
        st64(ses + 24, now);
        st64(ses + 40, ld64(ses + 40) + 1);
        st64(ses + 56, ld64(ses + 56) + len);
    } else {
        int have = 0;
        int parity = (fl & 0x20) != 0;
        if (fl & 1) { /* get_eim_mapping(), :469-528; eim_mapping: ext_ip@0 ext_port@4 created@8 last_used@16 ref@24 */
            u8 ek[8];
            eim_key(ek, sip, sp, proto);
            u8 *m = ora_rt_lookup(mp[M_EIM], ek);
            if (m) {
                st64(m + 16, now);
                st32(m, 24, ld32(m, 24) + 1);
                ++*stat(M_NAT_STATS, NS_EIM_HIT);
            } else {
                u16 ext = take_port(sub, parity, sp, sip, proto);
                if (!ext) {
                    ++*stat(M_NAT_STATS, NS_EXHAUST);
                } else {
                    u8 nm[32];
                    memset(nm, 0, 32);
                    st32(nm, 0, ld32(sub, 0));
                    st16(nm, 4, ext);
                    st64(nm + 8, now);
                    st64(nm + 16, now);
                    st32(nm, 24, 1);
                    if (ora_rt_update(mp[M_EIM], ek, nm, ORA_NOEXIST) == 0) {
                        ++*stat(M_NAT_STATS, NS_EIM_MISS);
                        m = ora_rt_lookup(mp[M_EIM], ek);
                    } else if ((m = ora_rt_lookup(mp[M_EIM], ek)) != NULL) {
                        st32(m, 24, ld32(m, 24) + 1);
                        ++*stat(M_NAT_STATS, NS_EIM_HIT);
                    }
                }
            }
            if (m) {
                nip = ld32(m, 0);
                nport = swab16(ld16(m, 4));
                have = 1;
            }
        }
        if (!have) {
            u16 ap = take_port(sub, parity, swab16(sp), sip, proto);
            if (!ap) {
                ++*stat(M_NAT_STATS, NS_EXHAUST);
                ++*stat(M_NAT_STATS, NS_DROPPED);
                natlog(5, ld32(sub, 24), sip, ld32(sub, 0), sp, 0, dip, dp, proto, 0);
                return 2;
            }
            nip = ld32(sub, 0);
            nport = swab16(ap);
        }
        u8 ns[80]; /* new session, :711-728 */
        memset(ns, 0, 80);
        st32(ns, 0, nip);
        st16(ns, 4, nport);
        st16(ns, 6, sp);
        st32(ns, 8, sip);
        st32(ns, 12, dip);
        st16(ns, 16, dp);
        st64(ns + 24, now);
        st64(ns + 32, now);
        st64(ns + 40, 1);
        st64(ns + 56, len);
        ns[73] = proto;
        ns[75] = hairpin;
        ora_rt_update(mp[M_SESSIONS], key, ns, ORA_ANY);
        u8 rk[16];
        memset(rk, 0, 16);
        st32(rk, 0, dip);
        st32(rk, 4, nip);
        st16(rk, 8, dp);
        st16(rk, 10, nport);
        rk[12] = proto;
        ora_rt_update(mp[M_REVERSE], rk, key, ORA_ANY);
        st64(sub + 32, ld64(sub + 32) + 1);
        st64(sub + 40, ld64(sub + 40) + 1);
        ++*stat(M_NAT_STATS, NS_CREATED);
        natlog(1, ld32(sub, 24), sip, nip, sp, nport, dip, dp, proto, hairpin);
    }
    /* SNAT rewrite, :752-798.  The reference works on the packet in place and reads the old port / id
     * back from memory AFTER it has stored the new address: with ihl < 5 the L4 header overlaps the IP
     * header and that read returns address bytes — the order of loads and stores below is the reference's. */
    st32(f, 26, nip);
    st16(f, 24, csum32(ld16(f, 24), sip, nip));
    if (proto == 6) {
        u16 op = ld16(f, l4);
        st16(f, l4, nport);
        st16(f, l4 + 16, csum32(ld16(f, l4 + 16), sip, nip));
        st16(f, l4 + 16, csum16(ld16(f, l4 + 16), op, nport));
    } else if (proto == 17) {
        u16 op = ld16(f, l4);
        st16(f, l4, nport);
        if (ld16(f, l4 + 6)) {
            st16(f, l4 + 6, csum32(ld16(f, l4 + 6), sip, nip));
            st16(f, l4 + 6, csum16(ld16(f, l4 + 6), op, nport));
            if (ld16(f, l4 + 6) == 0) st16(f, l4 + 6, 0xffff);
        }
    } else {
        u16 oid = ld16(f, l4 + 4);
        st16(f, l4 + 4, nport);
        st16(f, l4 + 2, csum16(ld16(f, l4 + 2), oid, nport));
    }
    ++*stat(M_NAT_STATS, NS_SNAT);
    return 0;
}

static int p_nat_ingress(ora_pkt *p) { /* nat44_ingress, :805-948 */
    u8 *f = p->data;
    u32 len = p->len;
    if (len < 14 || ld16(f, 12) != 0x0008 || len < 34) return 0;
    u32 sip = ld32(f, 26), dip = ld32(f, 30);
    u8 proto = f[23];
    u32 l4 = 14 + (f[14] & 15) * 4;
    u16 sp = 0, dp = 0;
    if (proto == 6 || proto == 17) {
        if (l4 + (proto == 6 ? 20 : 8) > len) return 0;
        sp = ld16(f, l4);
        dp = ld16(f, l4 + 2);
    } else if (proto == 1) {
        if (l4 + 8 > len) return 0;
        dp = ld16(f, l4 + 4);
    } else {
        return 0;
    }
    u8 rk[16];
    memset(rk, 0, 16);
    st32(rk, 0, sip);
    st32(rk, 4, dip);
    st16(rk, 8, sp);
    st16(rk, 10, dp);
    rk[12] = proto;
    u8 *ok = ora_rt_lookup(mp[M_REVERSE], rk);
    if (!ok) {
        ++*stat(M_NAT_STATS, NS_PASSED);
        return 0;
    }
    u8 *ses = ora_rt_lookup(mp[M_SESSIONS], ok);
    if (!ses) {
        ora_rt_delete(mp[M_REVERSE], rk);
        ++*stat(M_NAT_STATS, NS_EXPIRED);
        return 0;
    }
    st64(ses + 24, ora_rt_now());
    st64(ses + 48, ld64(ses + 48) + 1);
    st64(ses + 64, ld64(ses + 64) + len);
    if (proto == 6) {
        u8 tf = f[l4 + 13];
        if (tf & 0x05)
            ses[72] = 3;
        else if (ses[72] == 0 && (tf & 0x10))
            ses[72] = 1;
    }
    u32 nip = ld32(ses, 8);
    u16 nport = ld16(ses, 6);
    st32(f, 30, nip); /* DNAT rewrite, :896-944: same in-place load/store order as the reference */
    st16(f, 24, csum32(ld16(f, 24), dip, nip));
    if (proto == 6) {
        u16 op = ld16(f, l4 + 2);
        st16(f, l4 + 2, nport);
        st16(f, l4 + 16, csum32(ld16(f, l4 + 16), dip, nip));
        st16(f, l4 + 16, csum16(ld16(f, l4 + 16), op, nport));
    } else if (proto == 17) {
        u16 op = ld16(f, l4 + 2);
        st16(f, l4 + 2, nport);
        if (ld16(f, l4 + 6)) {
            st16(f, l4 + 6, csum32(ld16(f, l4 + 6), dip, nip));
            st16(f, l4 + 6, csum16(ld16(f, l4 + 6), op, nport));
            if (ld16(f, l4 + 6) == 0) st16(f, l4 + 6, 0xffff);
        }
    } else {
        u16 oid = ld16(f, l4 + 4);
        st16(f, l4 + 4, nport);
        st16(f, l4 + 2, csum16(ld16(f, l4 + 2), oid, nport));
    }
    ++*stat(M_NAT_STATS, NS_DNAT);
    return 0;
}

static int p_nat_hairpin_xdp(ora_pkt *p) { /* nat44_hairpin_xdp, :951-991 */
    const u8 *f = p->data;
    u32 zero = 0;
    const u8 *cfg = ora_rt_lookup(mp[M_NAT_CFG], &zero);
    if (!cfg || !(ld32(cfg, 0) & 4)) return 2;
    if (p->len < 14 || ld16(f, 12) != 0x0008 || p->len < 34) return 2;
    u32 sip = ld32(f, 26), dip = ld32(f, 30);
    if (private_src(sip) && ora_rt_lookup(mp[M_HAIRPIN], &dip)) ++*stat(M_NAT_STATS, NS_HAIRPIN);
    return 2;
}

/* ======================= dhcp_fastpath_prog, bpf/dhcp_fastpath.c:619-813 ======================= */
enum { DS_TOTAL, DS_HIT, DS_MISS, DS_ERROR, DS_EXPIRED, DS_O82, DS_O82_ABSENT, DS_BCAST, DS_UCAST, DS_VLAN };

static int p_dhcp(ora_pkt *p) {
    u8 *f = p->data;
    u32 len = p->len;
    if (len < 14) return 2;
    u16 et = ld16(f, 12);
    u32 l3 = 14, vo = 0, s_tag = 0, c_tag = 0;
    int tagged = 0;
    if (et == 0x0081 || et == 0xA888) { /* parse_packet_headers(), :352-428 */
        if (len < 18) return 2;
        tagged = 1;
        s_tag = swab16(ld16(f, 14)) & 0x0fff;
        vo = 4;
        et = ld16(f, 16);
        l3 = 18;
        ++*stat(M_DHCP_STATS, DS_VLAN);
        if (et == 0x0081) {
            if (len < 22) return 2;
            c_tag = swab16(ld16(f, 18)) & 0x0fff;
            vo = 8;
            et = ld16(f, 20);
            l3 = 22;
        }
    }
    if (et != 0x0008 || l3 + 20 > len || f[l3 + 9] != 17) return 2;
    u32 udp = l3 + (f[l3] & 15) * 4;
    if (udp + 8 > len || ld16(f, udp + 2) != 0x4300) return 2;
    u32 dh = udp + 8;
    if (dh + 240 > len) return 2;
    if (f[dh] != 1 || ld32(f, dh + 236) != 0x63538263) return 2;
    ++*stat(M_DHCP_STATS, DS_TOTAL);
    u32 op = dh + 240;
    u8 mt = 0;
    if (op + 12 <= len) { /* get_dhcp_msg_type(), :216-250 */
        static const u8 at[6] = {0, 1, 3, 4, 5, 6};
        for (int i = 0; i < 6 && !mt; i++)
            if (f[op + at[i]] == 53 && f[op + at[i] + 1] == 1) {
                mt = f[op + at[i] + 2];
                break;
            }
    }
    if (mt != 1 && mt != 3) {
        ++*stat(M_DHCP_STATS, DS_MISS);
        return 2;
    }
    const u8 *a = NULL; /* pool_assignment: pool_id@0 allocated_ip@4 lease_expiry@13 */
    if (tagged) {
        u8 vk[4];
        st16(vk, 0, (u16)s_tag);
        st16(vk, 2, (u16)c_tag);
        a = ora_rt_lookup(mp[M_VLANPOOLS], vk);
    }
    if (!a && op + 64 <= len) { /* extract_circuit_id_fixed(), :267-323 */
        const u8 *o = f + op;
        int at = -1;
        u32 cl = 0;
        if (o[3] == 82 && o[4] >= 4 && op + 5 + o[4] <= len && o[5] == 1 && o[6] > 0 && o[6] <= 32 && op + 7 + o[6] <= len) {
            at = 7;
            cl = o[6];
        }
        for (int q = 12; q < 20 && at < 0; q++)
            if (o[q] == 82 && op + q + 8 <= len && o[q + 1] >= 4 && o[q + 2] == 1 && o[q + 3] > 0 && o[q + 3] <= 32 &&
                op + q + 4 + o[q + 3] <= len) {
                at = q + 4;
                cl = o[q + 3];
            }
        if (at >= 0) {
            u8 ck[32];
            memset(ck, 0, 32);
            memcpy(ck, o + at, cl);
            a = ora_rt_lookup(mp[M_CIDSUBS], ck);
            if (a) ++*stat(M_DHCP_STATS, DS_O82);
        }
    }
    if (!a) {
        u64 mk = mac48(f + dh + 28);
        a = ora_rt_lookup(mp[M_SUBPOOLS], &mk);
    }
    if (!a) {
        ++*stat(M_DHCP_STATS, DS_MISS);
        return 2;
    }
    if (ora_rt_now() / 1000000000ULL > ld64(a + 13)) {
        ++*stat(M_DHCP_STATS, DS_EXPIRED);
        return 2;
    }
    u32 pid = ld32(a, 0);
    const u8 *pool = ora_rt_lookup(mp[M_IPPOOLS], &pid); /* ip_pool: prefix@4 gateway@8 dns1@12 dns2@16 lease@20 */
    if (!pool) {
        ++*stat(M_DHCP_STATS, DS_ERROR);
        return 2;
    }
    ++*stat(M_DHCP_STATS, DS_HIT);
    u32 zero = 0;
    const u8 *cfg = ora_rt_lookup(mp[M_SRV_CFG], &zero); /* server_mac@0 server_ip@8 */
    if (!cfg) {
        ++*stat(M_DHCP_STATS, DS_ERROR);
        return 2;
    }
    u32 gw = ld32(pool, 8), srv = ld32(cfg, 8) ? ld32(cfg, 8) : gw, gi = ld32(f, dh + 24);
    if (gi) { /* relayed, :726-743 */
        memcpy(f, f + 6, 6);
        st32(f, l3 + 16, gi);
        st16(f, udp + 2, 0x4300);
        ++*stat(M_DHCP_STATS, DS_UCAST);
    } else { /* setup_reply_l2_headers(), :436-482 */
        if ((swab16(ld16(f, dh + 10)) & 0x8000) || ld32(f, dh + 12) == 0) {
            memset(f, 0xff, 6);
            ++*stat(M_DHCP_STATS, DS_BCAST);
        } else {
            memcpy(f, f + dh + 28, 6);
            ++*stat(M_DHCP_STATS, DS_UCAST);
        }
        st32(f, l3 + 16, 0xffffffff);
        st16(f, udp + 2, 0x4400);
    }
    memcpy(f + 6, cfg, 6);
    st32(f, l3 + 12, srv);
    f[l3 + 8] = 64;
    st16(f, l3 + 10, 0);
    st16(f, udp, 0x4300);
    st16(f, udp + 6, 0);
    f[dh] = 2;
    f[dh + 3] = 0;
    st32(f, dh + 16, ld32(a, 4));
    st32(f, dh + 20, srv);
    memset(f + dh + 44, 0, 192);
    if (op + 64 > len) return 2; /* options bounds check AFTER the rewrite, :769 */
    u8 *o = f + op; /* build_dhcp_options(), :519-602 */
    u32 n = 0, lease = ld32(pool, 20), pl = pool[4];
    u32 mask = pl == 0 ? 0 : (pl >= 32 ? 0xffffffff : swab32(0xffffffffu << (32 - pl)));
    o[n++] = 53, o[n++] = 1, o[n++] = mt == 1 ? 2 : 5;
#define OPT32(code, v) (o[n++] = (code), o[n++] = 4, st32(o, n, (v)), n += 4)
    OPT32(54, srv);
    OPT32(51, swab32(lease));
    OPT32(1, mask);
    OPT32(3, gw);
    if (ld32(pool, 12)) {
        int two = ld32(pool, 16) != 0;
        o[n++] = 6, o[n++] = two ? 8 : 4;
        st32(o, n, ld32(pool, 12)), n += 4;
        if (two) st32(o, n, ld32(pool, 16)), n += 4;
    }
    OPT32(58, swab32(lease / 2));
    OPT32(59, swab32((lease * 7u) / 8u));
    o[n++] = 255;
    u16 udp_len = (u16)(8 + 240 + n), ip_len = (u16)(20 + udp_len), total = (u16)(14 + vo + ip_len);
    st16(f, l3 + 2, swab16(ip_len));
    st16(f, udp + 4, swab16(udp_len));
    u32 s = 0; /* ip_checksum(), :488-503 */
    for (int i = 0; i < 10; i++) s += ld16(f, l3 + 2 * i);
    st16(f, l3 + 10, fold(s));
    int delta = (int)total - (int)(u16)len;
    if (delta) {
        if ((long)len + delta < 14) { /* bpf_xdp_adjust_tail() refuses */
            ++*stat(M_DHCP_STATS, DS_ERROR);
            return 2;
        }
        p->len = len + delta;
    }
    return 3;
}

/* ---- plumbing ---- */
static const ora_prog_desc k_progs[] = {
    {"antispoof_ingress", p_antispoof}, {"qos_egress_prog", p_qos_egress},     {"qos_ingress_prog", p_qos_ingress},
    {"nat44_egress", p_nat_egress},     {"nat44_ingress", p_nat_ingress},      {"nat44_hairpin_xdp", p_nat_hairpin_xdp},
    {"dhcp_fastpath_prog", p_dhcp},
};
const char *ora_impl_name(void) { return "port"; }
const ora_map_desc *ora_impl_maps(int *n) {
    *n = M_COUNT;
    return k_maps;
}
const ora_prog_desc *ora_impl_progs(int *n) {
    *n = (int)(sizeof(k_progs) / sizeof(k_progs[0]));
    return k_progs;
}
void ora_impl_bind(void) {
    for (int i = 0; i < M_COUNT; i++) mp[i] = ora_rt_map(i);
}

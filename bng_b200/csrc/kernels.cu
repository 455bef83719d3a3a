// bng_b200 — batch kernels for the TC programs (antispoof, QoS, NAT44) and the
// upstream pipeline.
//
// Execution model.  Frames of a batch must take effect in index
// order.  Three phases run back to back on the context's stream:
//   CLASSIFY  one thread per frame over a persistent grid (a multiple of the
//             SM count).  The first 64 bytes of the frame are held in
//             registers (4 x 128-bit loads), the first probe slot of every
//             table the frame may need is fetched up front so the loads are in
//             flight together, and everything whose effect commutes is
//             finished here: verdicts, session counters (atomics), the in-place
//             SNAT rewrite (128-bit stores of the touched chunks).  Frames
//             whose effect depends on earlier frames of the same subscriber
//             (new NAT flows, token-bucket decisions) only get an ordering key.
//   GROUP     a stable LSD radix sort (8-bit digits) of (key, frame index)
//             pairs groups those frames by subscriber while keeping index
//             order inside a group.  The first pass also compacts: frames
//             without a key never enter the sort, and every later kernel reads
//             its element count from device memory, so an all-hit batch costs
//             a few empty launches and no host round trip.
//   RESOLVE   one warp per subscriber walks its group 32 frames at a time:
//             lanes gather the frames' lengths in parallel, new flows are
//             created in index order, and the token bucket is applied with
//             warp-uniform fast paths (whole chunk passes / whole chunk drops)
//             before falling back to a lane-by-lane scan.
#include <string.h>

#include "kernels.h"
#include "progs.cuh"

#define BLOCK 256
#define MISS_FLAG 0x80000000u

// device-side counters (Scratch::counters)
enum { CNT_M = 0, CNT_NSEG = 1, CNT_DEFERRED = 2 };

// ---------------------------------------------------------------------------
// antispoof_ingress
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(BLOCK) k_antispoof(const __grid_constant__ DevCtx c, const __grid_constant__ DevBatch b) {
    __shared__ BlockStats bs;
    bstats_init(bs);
    u32 cfg = *(const u16 *)c.as_config;
    u32 n_allowed = 0;
    for (u32 i = blockIdx.x * BLOCK + threadIdx.x; i < b.n; i += gridDim.x * BLOCK) {
        u32 len = b.len[i];
        const u8 *p = frame_ptr(b, i);
        Hdr64 h;
        hdr_load(h, p, len);
        u64 mk = mac_key(h, 6);
        const u8 *bind = len >= 14 ? tbl_find<1, false>(c.bindings, &mk) : nullptr;
        b.verdict[i] = (u8)antispoof_eval(c, bs, h, len, i, b.now, bind, cfg, n_allowed);
    }
    warp_stat_flush(bs, ST_AS_ALLOWED, n_allowed);
    bstats_flush(bs, c.stats);
}

// ---------------------------------------------------------------------------
// qos_egress_prog / qos_ingress_prog: classify
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(BLOCK)
    k_qos_classify(const __grid_constant__ DevCtx c, const __grid_constant__ DevBatch b, int egress, u32 *skey, u32 *sval) {
    __shared__ BlockStats bs;
    bstats_init(bs);
    const Tbl &t = egress ? c.qos_eg : c.qos_in;
    for (u32 i = blockIdx.x * BLOCK + threadIdx.x; i < b.n; i += gridDim.x * BLOCK) {
        u32 len = b.len[i];
        const u8 *p = frame_ptr(b, i);
        Hdr64 h;
        hdr_load(h, p, len < 34 ? len : 34);
        u32 prio;
        bool prio_set;
        u32 key = qos_classify_one(c, bs, t, h, len, egress != 0, &prio, &prio_set);
        if (prio_set && b.priority) b.priority[i] = prio;
        b.verdict[i] = TC_OK;
        skey[i] = key;
        sval[i] = i;
    }
    bstats_flush(bs, c.stats);
}

// ---------------------------------------------------------------------------
// nat44_egress: classify (generic path: any ihl, fields read from memory)
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(BLOCK)
    k_nat_eg_classify(const __grid_constant__ DevCtx c, const __grid_constant__ DevBatch b, u32 *skey, u32 *sval) {
    __shared__ BlockStats bs;
    bstats_init(bs);
    for (u32 i = blockIdx.x * BLOCK + threadIdx.x; i < b.n; i += gridDim.x * BLOCK) {
        u32 len = b.len[i];
        u8 *p = frame_ptr(b, i);
        NatOut o = nat_egress_one<false>(c, bs, p, len, i, b.now);
        b.verdict[i] = (u8)o.verdict;
        skey[i] = o.order_key;
        sval[i] = i | MISS_FLAG;
    }
    bstats_flush(bs, c.stats);
}

// ---------------------------------------------------------------------------
// nat44_ingress, nat44_hairpin_xdp
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(BLOCK) k_nat_ingress(const __grid_constant__ DevCtx c, const __grid_constant__ DevBatch b) {
    __shared__ BlockStats bs;
    bstats_init(bs);
    for (u32 i = blockIdx.x * BLOCK + threadIdx.x; i < b.n; i += gridDim.x * BLOCK)
        b.verdict[i] = (u8)nat_ingress_one(c, bs, frame_ptr(b, i), b.len[i], b.now);
    bstats_flush(bs, c.stats);
}

// bpf/nat44.c:951-991: statistics only, always XDP_PASS (2).
__global__ void __launch_bounds__(BLOCK) k_nat_hairpin_xdp(const __grid_constant__ DevCtx c, const __grid_constant__ DevBatch b) {
    __shared__ BlockStats bs;
    bstats_init(bs);
    u32 flags = *(const u32 *)c.nat_config;
    for (u32 i = blockIdx.x * BLOCK + threadIdx.x; i < b.n; i += gridDim.x * BLOCK) {
        b.verdict[i] = 2;
        if (!(flags & NATF_HAIRPIN)) continue;
        u32 len = b.len[i];
        const u8 *p = frame_ptr(b, i);
        if (len < 14 || rd16(p, 12) != ETH_P_IP_LE || len < 34) continue;
        if (!is_private_ip(rd32(p, 26))) continue;
        u64 hk = rd32(p, 30);
        if (tbl_find<1, false>(c.hairpin, &hk)) bstats_add(bs, ST_NAT_HAIRPIN, 1);
    }
    bstats_flush(bs, c.stats);
}

// ---------------------------------------------------------------------------
// pipeline_up classify: antispoof_ingress -> nat44_egress -> qos_ingress_prog
// keyed on the pre-NAT source address.  All three stages
// key their mutable state on the subscriber's private address, so one
// group-by serves both the NAT new-flow ordering and the token-bucket
// ordering.  Ordering key: the qos_ingress bucket slot when the subscriber
// has a bucket, else qos capacity + subscriber_nat slot.  MISS_FLAG in the
// value marks frames whose NAT session has to be created in the ordered phase.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(BLOCK, 6)
    k_pipe_classify(const __grid_constant__ DevCtx c, const __grid_constant__ DevBatch b, u32 *skey, u32 *sval) {
    __shared__ SmallTabs st;
    __shared__ BlockStats bs;
    __shared__ u64 bar;
    smem_stage_begin(&st, c.small, (u32)sizeof(SmallTabs), &bar);
    bstats_init(bs);
    smem_stage_wait(&bar);
    const u32 as_cfg = st.as_cfg, nflags = st.nat_flags;
    u32 n_allowed = 0, n_snat = 0;
    for (u32 i = blockIdx.x * BLOCK + threadIdx.x; i < b.n; i += gridDim.x * BLOCK) {
        u32 len = b.len[i];
        u8 *p = frame_ptr(b, i);
        Hdr64 h;
        hdr_load(h, p, len);

        // ---- issue the first probe of every table this frame may need ----
        const bool ip4 = len >= 34 && h.b16(12) == ETH_P_IP_LE;
        const u32 saddr = h.b32(26), daddr = h.b32(30), proto = h.b8(23);
        const bool ihl5 = (h.b8(14) & 0x0f) == 5;
        u64 mk = mac_key(h, 6);
        u32 bi = tbl_hash<1>(&mk) & c.bindings.mask;
        u64 bw0 = *(const u64 *)tbl_slot(c.bindings, bi);
        u64 sk = saddr;
        u32 ai = tbl_hash<1>(&sk); // subscriber_nat and qos_ingress share the key, hence the hash
        u32 si = ai & c.sub_nat.mask, qi = ai & c.qos_in.mask;
        u16 sport = 0, dport = 0;
        if (proto == 1) {
            sport = h.b16(38);
        } else {
            sport = h.b16(34);
            dport = h.b16(36);
        }
        u64 key[2];
        key[0] = (u64)saddr | ((u64)daddr << 32);
        key[1] = (u64)sport | ((u64)dport << 16) | ((u64)proto << 32);
        u32 hi = tbl_hash<2>(key) & c.sessions.mask;
        u64 sw0 = K_EMPTY, qw0 = K_EMPTY, kw0 = K_EMPTY, kw1 = 0;
        if (ip4) {
            sw0 = *(const u64 *)tbl_slot(c.sub_nat, si);
            qw0 = *(const u64 *)tbl_slot(c.qos_in, qi);
            const ulonglong2 kk = *(const ulonglong2 *)tbl_slot(c.sessions, hi);
            kw0 = kk.x;
            kw1 = kk.y;
        }

        // ---- antispoof_ingress ----
        const u8 *bind = len >= 14 ? tbl_finish<1>(c.bindings, &mk, bi, bw0, true) : nullptr;
        int v = antispoof_eval(c, bs, h, len, i, b.now, bind, as_cfg, n_allowed);
        u32 okey = NO_KEY, oval = i;
        if (v != TC_SHOT && ip4) {
            // ---- qos_ingress bucket of the pre-NAT source (nothing counted yet) ----
            const u8 *qsl = tbl_finish<1>(c.qos_in, &sk, qi, qw0, true);
            u32 qs = qsl ? (u32)((qsl - c.qos_in.slots) / c.qos_in.slot_bytes) : NO_KEY;
            // ---- nat44_egress ----
            bool miss = false;
            u32 sub_idx = 0;
            if (!ihl5) { // IPv4 options: fields are not at fixed offsets, take the generic path
                NatOut o = nat_egress_one<false>(c, bs, p, len, i, b.now);
                v = o.verdict;
                miss = o.order_key != NO_KEY;
                sub_idx = o.order_key;
            } else if (is_private_ip(saddr)) {
                const u8 *sub = tbl_finish<1>(c.sub_nat, &sk, si, sw0, true);
                if (!sub) {
                    bstats_add(bs, ST_NAT_PASSED, 1);
                } else {
                    bool stop = false;
                    if (proto == 6 || proto == 17) {
                        if ((proto == 6 ? 54u : 42u) > len) {
                            stop = true;
                        } else if ((nflags & (proto == 6 ? (NATF_ALG_FTP | NATF_ALG_SIP) : NATF_ALG_SIP)) && st.alg_n) {
                            int ai2 = alg_find(st, ((u32)bswap16(dport) << 16) | proto);
                            if (ai2 >= 0) { // bpf/nat44.c:615-642
                                bstats_add(bs, ST_NAT_ALG, 1);
                                nat_log(c, i, b.now, 7, *(const u32 *)(sub + 32), saddr, 0, sport, 0, daddr, dport, (u8)proto,
                                        st.alg_type[ai2]);
                                stop = true;
                            }
                        }
                    } else if (proto == 1) {
                        if (42u > len) stop = true;
                    } else {
                        stop = true;
                    }
                    if (!stop) {
                        if ((nflags & NATF_HAIRPIN) && hp_contains(st, daddr)) bstats_add(bs, ST_NAT_HAIRPIN, 1);
                        u8 *ses = tbl_finish<2>(c.sessions, key, hi, kw0, kw1 == key[1]);
                        if (ses) { // bpf/nat44.c:674-680 and the rewrite of :752-798, on the register copy
                            u32 nat_ip = *(const u32 *)(ses + 16);
                            u16 nat_port = *(const u16 *)(ses + 20);
                            *(u64 *)(ses + 40) = b.now;
                            atomicAdd((u64 *)(ses + 56), 1ull);
                            atomicAdd((u64 *)(ses + 72), (u64)len);
                            h.s32(26, nat_ip);
                            h.s16(24, csum_upd32(h.b16(24), saddr, nat_ip));
                            if (proto == 6) {
                                h.s16(34, nat_port);
                                u16 ck = csum_upd32(h.b16(50), saddr, nat_ip);
                                h.s16(50, csum_upd16(ck, sport, nat_port));
                                hdr_store_chunk(h, p, 3);
                            } else if (proto == 17) {
                                h.s16(34, nat_port);
                                u16 ck = h.b16(40);
                                if (ck != 0) {
                                    ck = csum_upd32(ck, saddr, nat_ip);
                                    ck = csum_upd16(ck, sport, nat_port);
                                    if (ck == 0) ck = 0xffff;
                                    h.s16(40, ck);
                                }
                            } else {
                                h.s16(38, nat_port);
                                h.s16(36, csum_upd16(h.b16(36), sport, nat_port));
                            }
                            hdr_store_chunk(h, p, 1);
                            hdr_store_chunk(h, p, 2);
                            n_snat++;
                        } else {
                            miss = true;
                            sub_idx = (u32)((sub - c.sub_nat.slots) / c.sub_nat.slot_bytes);
                        }
                    }
                }
            }
            if (v != TC_SHOT) {
                if (qs != NO_KEY) {
                    u64 rate = *(const u64 *)(qsl + 32);
                    if (!miss && rate == 0) { // unlimited bucket and nothing left to order
                        bstats_add(bs, ST_QOS_PASS_PKTS, 1);
                        bstats_add(bs, ST_QOS_PASS_BYTES, len);
                    } else {
                        okey = qs;
                    }
                } else if (miss) {
                    okey = (c.qos_in.mask + 1) + sub_idx;
                }
                if (miss) oval |= MISS_FLAG;
            }
        }
        b.verdict[i] = (u8)v;
        skey[i] = okey;
        sval[i] = oval;
    }
    warp_stat_flush(bs, ST_AS_ALLOWED, n_allowed);
    warp_stat_flush(bs, ST_NAT_SNAT, n_snat);
    bstats_flush(bs, c.stats);
}

// ---------------------------------------------------------------------------
// GROUP: stable LSD radix sort with compaction and device-side counts.
// Pass 0 reads n (host-known) keys and skips NO_KEY entries; later passes
// read *cnt[CNT_M] elements.  Every block owns one contiguous range of the
// input, so (digit, block) order of the scanned histogram is index order.
// ---------------------------------------------------------------------------
#define RS_BLOCKS_PER_SM 4

__device__ __forceinline__ void rs_range(u32 total, u32 &lo, u32 &hi) {
    u32 per = (total + gridDim.x - 1) / gridDim.x;
    per = (per + BLOCK - 1) / BLOCK * BLOCK;
    u64 l = (u64)blockIdx.x * per, r = l + per;
    lo = l < total ? (u32)l : total;
    hi = r < total ? (u32)r : total;
}

__global__ void __launch_bounds__(BLOCK) k_rs_hist(const u32 *keys, u32 n_host, const u32 *cnt, int first, int shift, u32 *H) {
    __shared__ u32 h[256];
    h[threadIdx.x] = 0;
    __syncthreads();
    u32 total = first ? n_host : cnt[CNT_M];
    u32 lo, hi;
    rs_range(total, lo, hi);
    for (u32 i = lo + threadIdx.x; i < hi; i += BLOCK) {
        u32 k = keys[i];
        if (k != NO_KEY) atomicAdd(&h[(k >> shift) & 0xff], 1u);
    }
    __syncthreads();
    H[threadIdx.x * gridDim.x + blockIdx.x] = h[threadIdx.x];
}

// exclusive scan of the 256 x nblocks histogram in (digit, block) order; single block of 1024 threads
__global__ void __launch_bounds__(1024) k_rs_scan(u32 *H, u32 nblocks, u32 *cnt, int first) {
    __shared__ u32 part[1024];
    u32 total = 256 * nblocks;
    u32 per = (total + 1023) / 1024;
    u32 lo = threadIdx.x * per, hi = lo + per < total ? lo + per : total;
    u32 s = 0;
    for (u32 i = lo; i < hi; i++) s += H[i];
    part[threadIdx.x] = s;
    __syncthreads();
    for (u32 off = 1; off < 1024; off <<= 1) { // Hillis-Steele inclusive scan
        u32 v = threadIdx.x >= off ? part[threadIdx.x - off] : 0;
        __syncthreads();
        part[threadIdx.x] += v;
        __syncthreads();
    }
    u32 run = part[threadIdx.x] - s;
    for (u32 i = lo; i < hi; i++) {
        u32 v = H[i];
        H[i] = run;
        run += v;
    }
    if (first && threadIdx.x == 1023) cnt[CNT_M] = part[1023];
}

__global__ void __launch_bounds__(BLOCK) k_rs_scatter(const u32 *keys, const u32 *vals, u32 *okeys, u32 *ovals, u32 n_host,
                                                      const u32 *cnt, int first, int shift, const u32 *H) {
    __shared__ u32 offs[256];
    __shared__ u32 wcnt[BLOCK / 32][256];
    offs[threadIdx.x] = H[threadIdx.x * gridDim.x + blockIdx.x];
    u32 total = first ? n_host : cnt[CNT_M];
    u32 lo, hi;
    rs_range(total, lo, hi);
    const u32 lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (u32 base = lo; base < hi; base += BLOCK) {
        u32 i = base + threadIdx.x;
        u32 k = NO_KEY, v = 0;
        if (i < hi) {
            k = keys[i];
            v = vals[i];
        }
        bool valid = k != NO_KEY;
        u32 d = valid ? ((k >> shift) & 0xff) : (256 + lane); // invalid lanes match nobody
#pragma unroll
        for (int r = 0; r < BLOCK / 32; r++) wcnt[r][threadIdx.x] = 0;
        __syncthreads();
        u32 peers = __match_any_sync(0xffffffffu, d);
        u32 rank = __popc(peers & ((1u << lane) - 1));
        if (valid && rank == 0) wcnt[w][d] = __popc(peers);
        __syncthreads();
        { // thread d: running offsets of digit d across the warps of this tile, in warp (= index) order
            u32 run = offs[threadIdx.x];
#pragma unroll
            for (int r = 0; r < BLOCK / 32; r++) {
                u32 t = wcnt[r][threadIdx.x];
                wcnt[r][threadIdx.x] = run;
                run += t;
            }
            offs[threadIdx.x] = run;
        }
        __syncthreads();
        if (valid) {
            u32 pos = wcnt[w][d] + rank;
            okeys[pos] = k;
            ovals[pos] = v;
        }
        __syncthreads();
    }
}

// group heads: positions where the sorted key changes (any order; groups are independent)
__global__ void __launch_bounds__(BLOCK) k_heads(const u32 *skey, u32 *seg, u32 *cnt) {
    u32 m = cnt[CNT_M];
    for (u32 j = blockIdx.x * BLOCK + threadIdx.x; j < m; j += gridDim.x * BLOCK) {
        if (j == 0 || skey[j - 1] != skey[j]) seg[atomicAdd(&cnt[CNT_NSEG], 1u)] = j;
    }
}

// ---------------------------------------------------------------------------
// RESOLVE: one warp per group, frames in index order.
//   NAT:  frames flagged MISS_FLAG run the full sequential nat44_egress
//         (session re-lookup, EIM, port allocation, inserts, log, rewrite).
//   QOS:  token_bucket_check() for every surviving frame of the group.
// ---------------------------------------------------------------------------
template <bool NAT, bool QOS, bool EGRESS>
__global__ void __launch_bounds__(BLOCK) k_resolve(const __grid_constant__ DevCtx c, const __grid_constant__ DevBatch b,
                                                   const u32 *skey, const u32 *sval, const u32 *seg, const u32 *cnt) {
    __shared__ BlockStats bs;
    bstats_init(bs);
    const Tbl &qt = EGRESS ? c.qos_eg : c.qos_in;
    const u32 lane = threadIdx.x & 31;
    const u32 warp = (blockIdx.x * BLOCK + threadIdx.x) >> 5, nwarps = (gridDim.x * BLOCK) >> 5;
    const u32 m = cnt[CNT_M], nseg = cnt[CNT_NSEG];
    u64 pp = 0, pb = 0, dp = 0, db = 0; // per-lane partial QoS counters
    for (u32 s = warp; s < nseg; s += nwarps) {
        const u32 start = seg[s];
        const u32 key = skey[start];
        const bool has_bucket = QOS && key <= qt.mask;
        u8 *slot = has_bucket ? tbl_slot(qt, key) : nullptr;
        TokenBucket tb;
        if (has_bucket) tb_load(tb, slot);
        for (u32 q0 = start;; q0 += 32) {
            const u32 q = q0 + lane;
            const bool valid = q < m && skey[q] == key;
            const u32 vmask = __ballot_sync(0xffffffffu, valid);
            if (!vmask) break;
            const u32 sv = valid ? sval[q] : 0;
            const u32 idx = sv & ~MISS_FLAG;
            const u32 len = valid ? b.len[idx] : 0;
            bool dropped = false;
            if (NAT) {
                u32 mm = __ballot_sync(0xffffffffu, valid && (sv & MISS_FLAG));
                while (mm) { // new flows of this subscriber, strictly in index order
                    u32 l = __ffs(mm) - 1;
                    mm &= mm - 1;
                    if (lane == l) {
                        NatOut o = nat_egress_one<true>(c, bs, frame_ptr(b, idx), len, idx, b.now);
                        if (o.verdict == TC_SHOT) {
                            b.verdict[idx] = TC_SHOT;
                            dropped = true;
                        }
                    }
                    __syncwarp();
                }
            }
            if (has_bucket) {
                const bool elig = valid && !dropped;
                u32 em = __ballot_sync(0xffffffffu, elig);
                if (em) {
                    bool pass = true;
                    if (tb.rate_bps != 0) { // rate 0: unlimited, bucket untouched (bpf/qos_ratelimit.c:77-78)
                        tb_refill(tb, b.now);
                        u32 mylen = elig ? len : 0;
                        u32 tot = __reduce_add_sync(0xffffffffu, mylen);
                        u32 mn = __reduce_min_sync(0xffffffffu, elig ? len : 0xffffffffu);
                        if (tb.tokens >= (u64)tot) { // the whole chunk fits
                            tb.tokens -= tot;
                        } else if (tb.tokens < (u64)mn) { // nothing in the chunk fits
                            pass = false;
                        } else { // mixed: frame by frame, uniform across the warp
                            pass = false;
                            while (em) {
                                u32 l = __ffs(em) - 1;
                                em &= em - 1;
                                u32 ll = __shfl_sync(0xffffffffu, len, l);
                                bool ok = tb.tokens >= (u64)ll;
                                if (ok) tb.tokens -= ll;
                                if (lane == l) pass = ok;
                            }
                        }
                    }
                    if (elig) {
                        if (pass) {
                            pp++;
                            pb += len;
                            if (EGRESS && b.priority) b.priority[idx] = tb.prio;
                        } else {
                            dp++;
                            db += len;
                            b.verdict[idx] = TC_SHOT;
                        }
                    }
                }
            }
            if (vmask != 0xffffffffu) break; // the group ended inside this chunk
        }
        if (has_bucket && lane == 0) {
            *(u64 *)(slot + 16) = tb.tokens;
            *(u64 *)(slot + 24) = tb.last_update;
        }
    }
    if (QOS) {
        warp_stat_flush64(bs, ST_QOS_PASS_PKTS, pp);
        warp_stat_flush64(bs, ST_QOS_PASS_BYTES, pb);
        warp_stat_flush64(bs, ST_QOS_DROP_PKTS, dp);
        warp_stat_flush64(bs, ST_QOS_DROP_BYTES, db);
    }
    bstats_flush(bs, c.stats);
}

// ---------------------------------------------------------------------------
// host-side launchers
// ---------------------------------------------------------------------------
static inline int grid_for(const Launcher &L, u32 n, int blocks_per_sm) {
    long want = ((long)n + BLOCK - 1) / BLOCK;
    long cap = (long)L.num_sms * blocks_per_sm;
    if (want < 1) want = 1;
    return (int)(want < cap ? want : cap);
}

static int bits_for(u64 max_key_exclusive) {
    int b = 1;
    while ((1ull << b) < max_key_exclusive) b++;
    return b;
}

size_t sort_temp_bytes(u32 n) { // histogram matrix: 256 digits x blocks
    (void)n;
    return (size_t)256 * 4096 * sizeof(u32);
}

void prof_begin(Launcher &L, const char *name) {
    if (!L.prof || L.npend >= 32) return;
    int acc = -1;
    for (int i = 0; i < L.nacc; i++)
        if (L.acc_name[i] == name || !strcmp(L.acc_name[i], name)) acc = i;
    if (acc < 0) {
        if (L.nacc >= 32) return;
        acc = L.nacc++;
        L.acc_name[acc] = name;
        L.acc_ms[acc] = 0;
        L.acc_n[acc] = 0;
    }
    ProfPending &p = L.pend[L.npend];
    p.acc = acc;
    if (cudaEventCreate(&p.a) != cudaSuccess || cudaEventCreate(&p.b) != cudaSuccess) return;
    cudaEventRecord(p.a, L.stream);
    L.npend++;
    L.prof = 2; // a begin is open
}
void prof_end(Launcher &L) {
    if (L.prof != 2) return;
    cudaEventRecord(L.pend[L.npend - 1].b, L.stream);
    L.prof = 1;
}
void prof_collect(Launcher &L) {
    for (int i = 0; i < L.npend; i++) {
        float ms = 0;
        if (cudaEventElapsedTime(&ms, L.pend[i].a, L.pend[i].b) == cudaSuccess) {
            L.acc_ms[L.pend[i].acc] += ms;
            L.acc_n[L.pend[i].acc]++;
        }
        cudaEventDestroy(L.pend[i].a);
        cudaEventDestroy(L.pend[i].b);
    }
    L.npend = 0;
}

#define LAUNCH(kern, n, bps, ...)                                      \
    do {                                                               \
        prof_begin(L, #kern);                                          \
        kern<<<grid_for(L, n, bps), BLOCK, 0, L.stream>>>(__VA_ARGS__); \
        prof_end(L);                                                   \
        L.launches++;                                                  \
    } while (0)

// Groups the (key, value) pairs in (key_a, val_a)[0..n) by key, stably, skipping NO_KEY.
// On return *sk / *sv name the buffers holding the grouped pairs; counters[CNT_M] holds their
// number and seg[0..counters[CNT_NSEG]) the group heads.
static cudaError_t group_by_key(Launcher &L, u32 n, u64 key_space, const u32 **sk, const u32 **sv) {
    Scratch &s = L.s;
    int passes = (bits_for(key_space) + 7) / 8;
    int rsb = L.num_sms * RS_BLOCKS_PER_SM;
    if (rsb > 4096) rsb = 4096;
    u32 *H = (u32 *)s.cub_tmp;
    cudaError_t e = cudaMemsetAsync(s.counters, 0, 64, L.stream);
    if (e != cudaSuccess) return e;
    u32 *ik = s.key_a, *iv = s.val_a, *ok = s.key_b, *ov = s.val_b;
    prof_begin(L, "group_by_key");
    for (int p = 0; p < passes; p++) {
        int first = p == 0;
        k_rs_hist<<<rsb, BLOCK, 0, L.stream>>>(ik, n, s.counters, first, 8 * p, H);
        k_rs_scan<<<1, 1024, 0, L.stream>>>(H, (u32)rsb, s.counters, first);
        k_rs_scatter<<<rsb, BLOCK, 0, L.stream>>>(ik, iv, ok, ov, n, s.counters, first, 8 * p, H);
        L.launches += 3;
        u32 *t = ik;
        ik = ok;
        ok = t;
        t = iv;
        iv = ov;
        ov = t;
    }
    k_heads<<<grid_for(L, n, 8), BLOCK, 0, L.stream>>>(ik, s.qslot, s.counters);
    L.launches++;
    prof_end(L);
    *sk = ik;
    *sv = iv;
    return cudaGetLastError();
}

cudaError_t run_antispoof(Launcher &L, const DevCtx &c, const DevBatch &b) {
    LAUNCH(k_antispoof, b.n, 8, c, b);
    return cudaGetLastError();
}

cudaError_t run_qos(Launcher &L, const DevCtx &c, const DevBatch &b, bool egress) {
    LAUNCH(k_qos_classify, b.n, 8, c, b, egress ? 1 : 0, L.s.key_a, L.s.val_a);
    const Tbl &t = egress ? c.qos_eg : c.qos_in;
    const u32 *sk, *sv;
    cudaError_t e = group_by_key(L, b.n, (u64)t.mask + 1, &sk, &sv);
    if (e != cudaSuccess) return e;
    if (egress)
        LAUNCH((k_resolve<false, true, true>), b.n, 4, c, b, sk, sv, L.s.qslot, L.s.counters);
    else
        LAUNCH((k_resolve<false, true, false>), b.n, 4, c, b, sk, sv, L.s.qslot, L.s.counters);
    return cudaGetLastError();
}

cudaError_t run_nat_egress(Launcher &L, const DevCtx &c, const DevBatch &b) {
    LAUNCH(k_nat_eg_classify, b.n, 6, c, b, L.s.key_a, L.s.val_a);
    const u32 *sk, *sv;
    cudaError_t e = group_by_key(L, b.n, (u64)c.sub_nat.mask + 1, &sk, &sv);
    if (e != cudaSuccess) return e;
    LAUNCH((k_resolve<true, false, false>), b.n, 4, c, b, sk, sv, L.s.qslot, L.s.counters);
    return cudaGetLastError();
}

cudaError_t run_nat_ingress(Launcher &L, const DevCtx &c, const DevBatch &b) {
    LAUNCH(k_nat_ingress, b.n, 6, c, b);
    return cudaGetLastError();
}

cudaError_t run_nat_hairpin_xdp(Launcher &L, const DevCtx &c, const DevBatch &b) {
    LAUNCH(k_nat_hairpin_xdp, b.n, 8, c, b);
    return cudaGetLastError();
}

cudaError_t run_pipeline_up(Launcher &L, const DevCtx &c, const DevBatch &b) {
    LAUNCH(k_pipe_classify, b.n, 6, c, b, L.s.key_a, L.s.val_a);
    u64 space = (u64)(c.qos_in.mask + 1) + (c.sub_nat.mask + 1);
    const u32 *sk, *sv;
    cudaError_t e = group_by_key(L, b.n, space, &sk, &sv);
    if (e != cudaSuccess) return e;
    LAUNCH((k_resolve<true, true, false>), b.n, 4, c, b, sk, sv, L.s.qslot, L.s.counters);
    return cudaGetLastError();
}

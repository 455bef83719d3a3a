// bng_b200 — batch kernels for the TC programs (antispoof, QoS, NAT44) and the
// upstream pipeline, plus the stable group-by-subscriber step between the
// classify and resolve phases.
//
// Execution model: frames of a batch must take effect in index
// order.  CLASSIFY kernels run one thread per frame over a persistent grid
// (a multiple of the SM count) and finish every frame whose effect commutes;
// frames whose effect depends on earlier frames of the same subscriber get an
// ordering key.  A stable radix sort groups those frames by key, keeping index
// order inside a group, and RESOLVE kernels walk each group sequentially.
#include <cub/device/device_radix_sort.cuh>

#include <string.h>

#include "kernels.h"
#include "progs.cuh"

#define BLOCK 256

// ---------------------------------------------------------------------------
// antispoof_ingress
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(BLOCK) k_antispoof(const __grid_constant__ DevCtx c, const __grid_constant__ DevBatch b) {
    __shared__ BlockStats bs;
    bstats_init(bs);
    for (u32 i = blockIdx.x * BLOCK + threadIdx.x; i < b.n; i += gridDim.x * BLOCK) {
        u32 len = b.len[i];
        const u8 *p = frame_ptr(b, i);
        Hdr64 h;
        hdr_load(h, p, len);
        b.verdict[i] = (u8)antispoof_one(c, bs, h, len, i, b.now);
    }
    bstats_flush(bs, c.stats);
}

// ---------------------------------------------------------------------------
// qos_egress_prog / qos_ingress_prog
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

// One thread per group head walks its group (frames sharing a bucket) in index order.
__global__ void __launch_bounds__(BLOCK) k_qos_resolve(const __grid_constant__ DevCtx c, const __grid_constant__ DevBatch b,
                                                       int egress, const u32 *skey, const u32 *sval) {
    __shared__ BlockStats bs;
    bstats_init(bs);
    const Tbl &t = egress ? c.qos_eg : c.qos_in;
    for (u32 j = blockIdx.x * BLOCK + threadIdx.x; j < b.n; j += gridDim.x * BLOCK) {
        u32 key = skey[j];
        if (key == NO_KEY) continue;
        if (j > 0 && skey[j - 1] == key) continue;
        u8 *slot = tbl_slot(t, key);
        TokenBucket tb;
        tb_load(tb, slot);
        u64 pp = 0, pb = 0, dp = 0, db = 0;
        for (u32 q = j; q < b.n && skey[q] == key; q++) {
            u32 idx = sval[q];
            u32 len = b.len[idx];
            if (tb_step(tb, b.now, len)) {
                pp++;
                pb += len;
                if (egress && b.priority) b.priority[idx] = tb.prio;
            } else {
                dp++;
                db += len;
                b.verdict[idx] = TC_SHOT;
            }
        }
        *(u64 *)(slot + 16) = tb.tokens;
        *(u64 *)(slot + 24) = tb.last_update;
        if (pp) {
            bstats_add(bs, ST_QOS_PASS_PKTS, pp);
            bstats_add(bs, ST_QOS_PASS_BYTES, pb);
        }
        if (dp) {
            bstats_add(bs, ST_QOS_DROP_PKTS, dp);
            bstats_add(bs, ST_QOS_DROP_BYTES, db);
        }
    }
    bstats_flush(bs, c.stats);
}

// ---------------------------------------------------------------------------
// nat44_egress
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(BLOCK)
    k_nat_eg_classify(const __grid_constant__ DevCtx c, const __grid_constant__ DevBatch b, u32 *skey, u32 *sval, u32 *n_deferred) {
    __shared__ BlockStats bs;
    __shared__ u32 s_def;
    if (threadIdx.x == 0) s_def = 0;
    bstats_init(bs);
    for (u32 i = blockIdx.x * BLOCK + threadIdx.x; i < b.n; i += gridDim.x * BLOCK) {
        u32 len = b.len[i];
        u8 *p = frame_ptr(b, i);
        NatOut o = nat_egress_one<false>(c, bs, p, len, i, b.now);
        b.verdict[i] = (u8)o.verdict;
        skey[i] = o.order_key;
        sval[i] = i;
        if (o.order_key != NO_KEY) atomicAdd(&s_def, 1u);
    }
    bstats_flush(bs, c.stats);
    if (threadIdx.x == 0 && s_def) atomicAdd(n_deferred, s_def);
}

__global__ void __launch_bounds__(BLOCK) k_nat_eg_resolve(const __grid_constant__ DevCtx c, const __grid_constant__ DevBatch b,
                                                          const u32 *skey, const u32 *sval, const u32 *n_deferred) {
    __shared__ BlockStats bs;
    bstats_init(bs);
    if (*n_deferred != 0) {
        for (u32 j = blockIdx.x * BLOCK + threadIdx.x; j < b.n; j += gridDim.x * BLOCK) {
            u32 key = skey[j];
            if (key == NO_KEY) continue;
            if (j > 0 && skey[j - 1] == key) continue;
            for (u32 q = j; q < b.n && skey[q] == key; q++) {
                u32 idx = sval[q];
                NatOut o = nat_egress_one<true>(c, bs, frame_ptr(b, idx), b.len[idx], idx, b.now);
                b.verdict[idx] = (u8)o.verdict;
            }
        }
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
// pipeline_up: antispoof_ingress -> nat44_egress -> qos_ingress_prog keyed on
// the pre-NAT source address.  All three stages key their
// mutable state on the subscriber's private address, so one group-by serves
// both the NAT new-flow ordering and the token-bucket ordering.
//   pflag bit0: NAT session miss (run the ordered NAT step first)
//   qslot:      qos_ingress bucket slot or NO_KEY
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(BLOCK) k_pipe_classify(const __grid_constant__ DevCtx c, const __grid_constant__ DevBatch b,
                                                         u32 *skey, u32 *sval, u32 *qslot, u8 *pflag) {
    __shared__ BlockStats bs;
    bstats_init(bs);
    for (u32 i = blockIdx.x * BLOCK + threadIdx.x; i < b.n; i += gridDim.x * BLOCK) {
        u32 len = b.len[i];
        u8 *p = frame_ptr(b, i);
        Hdr64 h;
        hdr_load(h, p, len);
        u32 key = NO_KEY, qs = NO_KEY;
        u8 pf = 0;
        int v = antispoof_one(c, bs, h, len, i, b.now);
        if (v != TC_SHOT) {
            // qos_ingress lookup on the pre-NAT header (nothing is counted yet:
            // a NAT drop later in the pipeline must leave QoS untouched)
            if (len >= 34 && h.b16(12) == ETH_P_IP_LE) {
                u64 k = h.b32(26);
                const u8 *qsl = tbl_find<1, false>(c.qos_in, &k);
                if (qsl) qs = (u32)((qsl - c.qos_in.slots) / c.qos_in.slot_bytes);
            }
            NatOut o = nat_egress_one<false>(c, bs, p, len, i, b.now);
            v = o.verdict;
            if (o.order_key != NO_KEY) pf = 1;
            if (qs != NO_KEY) {
                u64 rate = *(const u64 *)(tbl_slot(c.qos_in, qs) + 32);
                if (pf == 0 && rate == 0) { // unlimited bucket and nothing left to order
                    bstats_add(bs, ST_QOS_PASS_PKTS, 1);
                    bstats_add(bs, ST_QOS_PASS_BYTES, len);
                } else {
                    key = qs;
                }
            } else if (pf) {
                key = (c.qos_in.mask + 1) + o.order_key;
            }
        }
        b.verdict[i] = (u8)v;
        skey[i] = key;
        sval[i] = i;
        qslot[i] = qs;
        pflag[i] = pf;
    }
    bstats_flush(bs, c.stats);
}

__global__ void __launch_bounds__(BLOCK) k_pipe_resolve(const __grid_constant__ DevCtx c, const __grid_constant__ DevBatch b,
                                                        const u32 *skey, const u32 *sval, const u32 *qslot, const u8 *pflag) {
    __shared__ BlockStats bs;
    bstats_init(bs);
    for (u32 j = blockIdx.x * BLOCK + threadIdx.x; j < b.n; j += gridDim.x * BLOCK) {
        u32 key = skey[j];
        if (key == NO_KEY) continue;
        if (j > 0 && skey[j - 1] == key) continue;
        bool has_bucket = key <= c.qos_in.mask;
        u8 *slot = has_bucket ? tbl_slot(c.qos_in, key) : nullptr;
        TokenBucket tb;
        if (has_bucket) tb_load(tb, slot);
        u64 pp = 0, pb = 0, dp = 0, db = 0;
        for (u32 q = j; q < b.n && skey[q] == key; q++) {
            u32 idx = sval[q];
            u32 len = b.len[idx];
            if (pflag[idx] & 1) {
                NatOut o = nat_egress_one<true>(c, bs, frame_ptr(b, idx), len, idx, b.now);
                if (o.verdict == TC_SHOT) {
                    b.verdict[idx] = TC_SHOT;
                    continue;
                }
            }
            if (!has_bucket) continue;
            if (tb.rate_bps == 0 || tb_step(tb, b.now, len)) {
                pp++;
                pb += len;
            } else {
                dp++;
                db += len;
                b.verdict[idx] = TC_SHOT;
            }
        }
        if (has_bucket) {
            *(u64 *)(slot + 16) = tb.tokens;
            *(u64 *)(slot + 24) = tb.last_update;
        }
        if (pp) {
            bstats_add(bs, ST_QOS_PASS_PKTS, pp);
            bstats_add(bs, ST_QOS_PASS_BYTES, pb);
        }
        if (dp) {
            bstats_add(bs, ST_QOS_DROP_PKTS, dp);
            bstats_add(bs, ST_QOS_DROP_BYTES, db);
        }
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

static int bits_for(u32 max_key_exclusive) {
    int b = 1;
    while ((1ull << b) < max_key_exclusive) b++;
    return b;
}

size_t sort_temp_bytes(u32 n) {
    size_t bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, bytes, (const u32 *)nullptr, (u32 *)nullptr, (const u32 *)nullptr,
                                    (u32 *)nullptr, (int)n, 0, 32);
    return bytes;
}

// stable sort of (key, frame index) pairs; NO_KEY sorts last because bit `bits` is set in it
static cudaError_t group_by_key(Launcher &L, u32 n, u32 key_space) {
    int end_bit = bits_for(key_space) + 1;
    if (end_bit > 32) end_bit = 32;
    size_t tb = L.s.cub_tmp_bytes;
    L.launches += (end_bit + 7) / 8 + 1;
    prof_begin(L, "group_by_key");
    cudaError_t e = cub::DeviceRadixSort::SortPairs(L.s.cub_tmp, tb, L.s.key_a, L.s.key_b, L.s.val_a, L.s.val_b, (int)n, 0,
                                                    end_bit, L.stream);
    prof_end(L);
    return e;
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

cudaError_t run_antispoof(Launcher &L, const DevCtx &c, const DevBatch &b) {
    LAUNCH(k_antispoof, b.n, 8, c, b);
    return cudaGetLastError();
}

cudaError_t run_qos(Launcher &L, const DevCtx &c, const DevBatch &b, bool egress) {
    LAUNCH(k_qos_classify, b.n, 8, c, b, egress ? 1 : 0, L.s.key_a, L.s.val_a);
    const Tbl &t = egress ? c.qos_eg : c.qos_in;
    cudaError_t e = group_by_key(L, b.n, t.mask + 1);
    if (e != cudaSuccess) return e;
    LAUNCH(k_qos_resolve, b.n, 8, c, b, egress ? 1 : 0, L.s.key_b, L.s.val_b);
    return cudaGetLastError();
}

cudaError_t run_nat_egress(Launcher &L, const DevCtx &c, const DevBatch &b) {
    cudaError_t e = cudaMemsetAsync(L.s.counters, 0, 64, L.stream);
    if (e != cudaSuccess) return e;
    LAUNCH(k_nat_eg_classify, b.n, 6, c, b, L.s.key_a, L.s.val_a, L.s.counters);
    e = group_by_key(L, b.n, c.sub_nat.mask + 1);
    if (e != cudaSuccess) return e;
    LAUNCH(k_nat_eg_resolve, b.n, 4, c, b, L.s.key_b, L.s.val_b, L.s.counters);
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
    LAUNCH(k_pipe_classify, b.n, 4, c, b, L.s.key_a, L.s.val_a, L.s.qslot, L.s.pflag);
    u64 space = (u64)(c.qos_in.mask + 1) + (c.sub_nat.mask + 1);
    cudaError_t e = group_by_key(L, b.n, (u32)space);
    if (e != cudaSuccess) return e;
    LAUNCH(k_pipe_resolve, b.n, 4, c, b, L.s.key_b, L.s.val_b, L.s.qslot, L.s.pflag);
    return cudaGetLastError();
}

// bng_b200 — dhcp_fastpath_prog as a batch kernel (bpf/dhcp_fastpath.c:619-813
// with bpf/maps.h).  One thread per frame: parse Eth/(QinQ)/IPv4/UDP/BOOTP,
// identify the subscriber (VLAN pair -> circuit-id -> chaddr), and rewrite the
// request into an OFFER/ACK in place.  No state depends on frame order, so the
// whole program is a single data-parallel pass; counters go through per-block
// shared accumulators.
#include "kernels.h"
#include "progs.cuh"

#define XDP_PASS_ 2
#define XDP_TX_ 3

// pool_assignment (packed, 25 B) at `a`: pool_id@0 allocated_ip@4 vlan_id@8
// client_class@12 lease_expiry@13 flags@21  (bpf/maps.h:89-97)
__device__ __forceinline__ u64 rd_u64_unaligned(const u8 *p) {
    u64 v = 0;
#pragma unroll
    for (int i = 7; i >= 0; i--) v = (v << 8) | p[i];
    return v;
}

__device__ __forceinline__ void zero_range(u8 *q, u32 n) { // q 2-byte aligned, n even
    while (n >= 2 && ((uintptr_t)q & 15)) {
        *(u16 *)q = 0;
        q += 2;
        n -= 2;
    }
    while (n >= 16) {
        *(uint4 *)q = make_uint4(0, 0, 0, 0);
        q += 16;
        n -= 16;
    }
    while (n >= 2) {
        *(u16 *)q = 0;
        q += 2;
        n -= 2;
    }
}

__device__ __forceinline__ void put_opt32(u8 *o, int &off, u8 code, u32 v_le) { // value bytes in memory order
    o[off++] = code;
    o[off++] = 4;
    o[off++] = (u8)v_le;
    o[off++] = (u8)(v_le >> 8);
    o[off++] = (u8)(v_le >> 16);
    o[off++] = (u8)(v_le >> 24);
}

// dlen: bytes of the frame present in its slot (every bounds check); len: the frame length (tail adjustment)
__device__ __forceinline__ int dhcp_one(const DevCtx &c, BlockStats &bs, u8 *p, u32 &len, const u32 dlen, u64 now) {
    // ---- parse_packet_headers(), :352-428 ----
    if (dlen < 14) return XDP_PASS_;
    u32 proto = rd16(p, 12);
    u32 l3 = 14, vlan_off = 0, vlan_id = 0, inner_id = 0;
    bool tagged = false;
    if (proto == 0x0081u || proto == 0xA888u) { // 802.1Q / 802.1ad
        if (dlen < 18) return XDP_PASS_;
        tagged = true;
        vlan_id = bswap16(rd16(p, 14)) & 0x0FFF;
        vlan_off = 4;
        proto = rd16(p, 16);
        l3 = 18;
        bstats_add(bs, ST_DHCP_VLAN, 1);
        if (proto == 0x0081u) {
            if (dlen < 22) return XDP_PASS_;
            inner_id = bswap16(rd16(p, 18)) & 0x0FFF;
            vlan_off = 8;
            proto = rd16(p, 20);
            l3 = 22;
        }
    }
    if (proto != ETH_P_IP_LE) return XDP_PASS_;
    if (l3 + 20 > dlen) return XDP_PASS_;
    if (p[l3 + 9] != 17) return XDP_PASS_;
    u32 udp = l3 + (u32)(p[l3] & 0x0f) * 4;
    if (udp + 8 > dlen) return XDP_PASS_;
    if (rd16(p, udp + 2) != 0x4300u) return XDP_PASS_; // bpf_htons(67)
    u32 dh = udp + 8;
    if (dh + 240 > dlen) return XDP_PASS_;

    // ---- :627-645 ----
    if (p[dh] != 1) return XDP_PASS_;
    if (rd32(p, dh + 236) != 0x63538263u) return XDP_PASS_; // bpf_htonl(0x63825363)
    bstats_add(bs, ST_DHCP_TOTAL, 1);
    u32 opts = dh + 240;
    u32 msg_type = 0;
    if (opts + 12 <= dlen) { // get_dhcp_msg_type(), :216-250
        const u8 *o = p + opts;
        if (o[0] == 53 && o[1] == 1) msg_type = o[2];
        else if (o[1] == 53 && o[2] == 1) msg_type = o[3];
        else if (o[3] == 53 && o[4] == 1) msg_type = o[5];
        else if (o[4] == 53 && o[5] == 1) msg_type = o[6];
        else if (o[5] == 53 && o[6] == 1) msg_type = o[7];
        else if (o[6] == 53 && o[7] == 1) msg_type = o[8];
    }
    if (msg_type != 1 && msg_type != 3) {
        bstats_add(bs, ST_DHCP_MISS, 1);
        return XDP_PASS_;
    }

    // ---- subscriber lookup, :653-687 ----
    const u8 *a = nullptr; // -> pool_assignment
    if (tagged) {
        u64 vk = (u64)(vlan_id | (inner_id << 16));
        const u8 *s = tbl_find_conv<1>(c.vlan_pools, &vk);
        if (s) a = s + c.vlan_pools.voff;
    }
    if (!a && opts + 64 <= dlen) { // extract_circuit_id_fixed(), :267-323
        const u8 *o = p + opts;
        int cid_at = -1;
        u32 cid_len = 0;
        if (o[3] == 82) {
            u32 l82 = o[4];
            if (l82 >= 4 && opts + 5 + l82 <= dlen && o[5] == 1) {
                u32 cl = o[6];
                if (cl > 0 && cl <= 32 && opts + 7 + cl <= dlen) {
                    cid_at = 7;
                    cid_len = cl;
                }
            }
        }
        if (cid_at < 0) {
            for (int pos = 12; pos < 20; pos++) {
                if (o[pos] == 82 && opts + pos + 8 <= dlen) {
                    u32 l82 = o[pos + 1];
                    if (l82 >= 4 && o[pos + 2] == 1) {
                        u32 cl = o[pos + 3];
                        if (cl > 0 && cl <= 32 && opts + pos + 4 + cl <= dlen) {
                            cid_at = pos + 4;
                            cid_len = cl;
                            break;
                        }
                    }
                }
            }
        }
        if (cid_at >= 0) {
            u64 ck[4] = {0, 0, 0, 0};
            for (u32 k = 0; k < cid_len; k++) ck[k >> 3] |= (u64)o[cid_at + k] << ((k & 7) * 8);
            const u8 *s = tbl_find_conv<4>(c.cid_subs, ck);
            if (s) {
                a = s + c.cid_subs.voff;
                bstats_add(bs, ST_DHCP_O82_PRESENT, 1);
            }
        }
    }
    if (!a) {
        u64 mk = 0;
#pragma unroll
        for (int k = 0; k < 6; k++) mk = (mk << 8) | p[dh + 28 + k]; // chaddr
        const u8 *s = tbl_find_conv<1>(c.sub_pools, &mk);
        if (s) a = s + c.sub_pools.voff;
    }
    if (!a) {
        bstats_add(bs, ST_DHCP_MISS, 1);
        return XDP_PASS_;
    }

    // ---- lease / pool / config, :689-713 ----
    u64 now_s = now / 1000000000ull;
    if (now_s > rd_u64_unaligned(a + 13)) {
        bstats_add(bs, ST_DHCP_EXPIRED, 1);
        return XDP_PASS_;
    }
    u64 pk = *(const u32 *)a;
    const u8 *pool = tbl_find_conv<1>(c.ip_pools, &pk);
    if (!pool) {
        bstats_add(bs, ST_DHCP_ERROR, 1);
        return XDP_PASS_;
    }
    pool += c.ip_pools.voff; // ip_pool: network@0 prefix_len@4 gateway@8 dns1@12 dns2@16 lease_time@20
    bstats_add(bs, ST_DHCP_HIT, 1);
    const u8 *cfg = c.server_config; // server_mac@0 server_ip@8
    u32 allocated_ip = *(const u32 *)(a + 4);
    u32 gateway = *(const u32 *)(pool + 8);
    u32 cfg_ip = *(const u32 *)(cfg + 8);
    u32 server_ip = cfg_ip != 0 ? cfg_ip : gateway;
    u8 reply_type = msg_type == 1 ? 2 : 5;
    u32 giaddr = rd32(p, dh + 24);
    u16 smac0 = *(const u16 *)(cfg + 0), smac1 = *(const u16 *)(cfg + 2), smac2 = *(const u16 *)(cfg + 4);

    if (giaddr != 0) { // relayed: unicast back to the relay agent, :726-743
        wr16(p, 0, rd16(p, 6));
        wr16(p, 2, rd16(p, 8));
        wr16(p, 4, rd16(p, 10));
        wr32(p, l3 + 16, giaddr);
        wr16(p, udp + 2, 0x4300u);
        bstats_add(bs, ST_DHCP_UCAST, 1);
    } else { // setup_reply_l2_headers(), :436-482
        u16 flags = bswap16(rd16(p, dh + 10));
        bool bcast = (flags & 0x8000) || rd32(p, dh + 12) == 0;
        if (bcast) {
            wr16(p, 0, 0xFFFF);
            wr16(p, 2, 0xFFFF);
            wr16(p, 4, 0xFFFF);
            bstats_add(bs, ST_DHCP_BCAST, 1);
        } else {
            wr16(p, 0, rd16(p, dh + 28));
            wr16(p, 2, rd16(p, dh + 30));
            wr16(p, 4, rd16(p, dh + 32));
            bstats_add(bs, ST_DHCP_UCAST, 1);
        }
        wr32(p, l3 + 16, 0xFFFFFFFFu);
        wr16(p, udp + 2, 0x4400u); // bpf_htons(68)
    }
    wr16(p, 6, smac0);
    wr16(p, 8, smac1);
    wr16(p, 10, smac2);
    wr32(p, l3 + 12, server_ip);
    p[l3 + 8] = 64;
    wr16(p, l3 + 10, 0);
    wr16(p, udp + 0, 0x4300u);
    wr16(p, udp + 6, 0);

    // ---- BOOTP fixed part, :758-766 ----
    p[dh + 0] = 2;
    p[dh + 3] = 0;
    wr32(p, dh + 16, allocated_ip);
    wr32(p, dh + 20, server_ip);
    zero_range(p + dh + 44, 192);

    // The options bounds check comes AFTER the header rewrite (:769): a frame
    // shorter than options+64 leaves here rewritten but XDP_PASSed.
    if (opts + 64 > dlen) return XDP_PASS_;

    // ---- build_dhcp_options(), :519-602 ----
    u8 *o = p + opts;
    int off = 0;
    u32 lease = *(const u32 *)(pool + 20);
    u32 plen = pool[4];
    u32 mask = plen == 0 ? 0u : (plen >= 32 ? 0xFFFFFFFFu : bswap32(0xFFFFFFFFu << (32 - plen)));
    o[off++] = 53;
    o[off++] = 1;
    o[off++] = reply_type;
    put_opt32(o, off, 54, server_ip);
    put_opt32(o, off, 51, bswap32(lease));
    put_opt32(o, off, 1, mask);
    put_opt32(o, off, 3, gateway);
    u32 dns1 = *(const u32 *)(pool + 12), dns2 = *(const u32 *)(pool + 16);
    if (dns1 != 0) {
        o[off++] = 6;
        o[off++] = dns2 != 0 ? 8 : 4;
        for (int k = 0; k < 4; k++) o[off++] = (u8)(dns1 >> (8 * k));
        if (dns2 != 0)
            for (int k = 0; k < 4; k++) o[off++] = (u8)(dns2 >> (8 * k));
    }
    put_opt32(o, off, 58, bswap32(lease / 2));
    put_opt32(o, off, 59, bswap32((lease * 7u) / 8u));
    o[off++] = 255;

    // ---- lengths, checksum, tail adjust, :779-812 ----
    u16 dhcp_len = (u16)(240 + off);
    u16 udp_len = (u16)(8 + dhcp_len);
    u16 ip_len = (u16)(20 + udp_len);
    u16 total = (u16)(14 + vlan_off + ip_len);
    wr16(p, l3 + 2, bswap16(ip_len));
    wr16(p, udp + 4, bswap16(udp_len));
    u32 sum = 0;
#pragma unroll
    for (int k = 0; k < 10; k++) sum += rd16(p, l3 + 2 * k); // ip_checksum(), :488-503 (check is 0 here)
    sum = (sum & 0xFFFF) + (sum >> 16);
    sum = (sum & 0xFFFF) + (sum >> 16);
    wr16(p, l3 + 10, (u16)~sum);
    int delta = (int)total - (int)(u16)len;
    if (delta != 0) {
        long nl = (long)len + delta;
        if (nl < 14) { // bpf_xdp_adjust_tail() refuses to go below the Ethernet header
            bstats_add(bs, ST_DHCP_ERROR, 1);
            return XDP_PASS_;
        }
        len = (u32)nl;
    }
    return XDP_TX_;
}

// Tile kernel (one mbarrier per block; per-thread barriers without any block-wide synchronisation were measured
// in round 2 and are SLOWER: 1.19 vs 0.96 ms per 2^22 requests): a request is up to ~350 bytes that the program reads sparsely and rewrites almost
// entirely (L2 headers, BOOTP fixed part, 192 zeroed bytes, options), so frames are staged through
// shared memory with the TMA: every thread bulk-copies its frame (cp.async.bulk, completion on an
// mbarrier), runs the program on the shared-memory copy, and bulk-stores it back.  HBM sees two
// streaming passes per frame instead of scattered 1-16 byte accesses.
#define DH_TILE 128
// Bytes staged per frame, also the slot stride in shared memory.  400 = 16 x 25: a multiple of 16 (bulk
// copies) whose word stride (100) spreads same-offset accesses of the 32 lanes over 8 banks; 384 or 448
// would put them on 1 or 2.  Frames whose options end beyond it (QinQ + IPv4 options) run in place.
#define DH_SLOT 400

// DHCP_TILE_TMA: in a fixed-stride arena (a receive ring: what a NIC fills) the 128 frames of a tile are one
// contiguous run, so the whole tile moves with ONE bulk copy each way instead of one per frame — the per-frame
// version issues 2 x 2^22 TMA operations per batch, ~32 cycles apart on every SM, and that, not HBM, is what
// bounded it (double-buffering the per-frame copies made it slower, 0.97 -> 1.43 ms: profiles/r02_notes.md).
#ifndef DHCP_TILE_TMA
#define DHCP_TILE_TMA 1
#endif
__global__ void __launch_bounds__(DH_TILE) k_dhcp_fastpath(const __grid_constant__ DevCtx c, const __grid_constant__ DevBatch b) {
    extern __shared__ __align__(128) u8 stage[]; // DH_TILE * DH_SLOT
    __shared__ BlockStats bs;
    __shared__ u64 bar, bar1;
    bstats_init(bs);
    const u32 bar_a = (u32)__cvta_generic_to_shared(&bar), bar1_a = (u32)__cvta_generic_to_shared(&bar1);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar_a), "r"(DH_TILE));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar1_a));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    // tile mode: fixed stride no larger than a staging slot, every frame's bytes inside its slot
    const bool tile_mode = DHCP_TILE_TMA && !b.off16 && b.stride <= DH_SLOT && b.cap == b.stride;
    const u32 sstride = tile_mode ? b.stride : DH_SLOT; // slot stride in shared memory
    u8 *mine = stage + (size_t)threadIdx.x * sstride;
    const u32 mine_a = (u32)__cvta_generic_to_shared(mine), stage_a = (u32)__cvta_generic_to_shared(stage);
    u32 phase = 0;
    for (u32 base = blockIdx.x * DH_TILE; base < b.n; base += gridDim.x * DH_TILE) {
        const u32 i = base + threadIdx.x;
        const bool act = i < b.n;
        u32 len = act ? b.len[i] : 0;
        u8 *g = act ? frame_ptr(b, i) : b.pkts;
        const u32 present = frame_dlen(b, len);
        u32 nbytes = ((present < DH_SLOT ? present : DH_SLOT) + 15u) & ~15u;
        if (nbytes > DH_SLOT) nbytes = DH_SLOT;
        const u32 tile_bytes = (b.n - base < (u32)DH_TILE ? b.n - base : (u32)DH_TILE) * b.stride;
        // ---- stage in ----
        if (tile_mode) {
            if (threadIdx.x == 0) {
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar1_a), "r"(tile_bytes) : "memory");
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(stage_a),
                             "l"(b.pkts + (size_t)base * b.stride), "r"(tile_bytes), "r"(bar1_a)
                             : "memory");
            }
        } else if (nbytes) {
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_a), "r"(nbytes) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(mine_a),
                         "l"(g), "r"(nbytes), "r"(bar_a)
                         : "memory");
        } else {
            asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar_a) : "memory");
        }
        u32 done = 0;
        while (!done) {
            asm volatile(
                "{\n\t.reg .pred p;\n\t"
                "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                "selp.u32 %0, 1, 0, p;\n\t}"
                : "=r"(done)
                : "r"(tile_mode ? bar1_a : bar_a), "r"(phase)
                : "memory");
        }
        phase ^= 1;
        // ---- the program, on the staged copy ----
        bool direct = false; // the program would reach past the staged bytes: run it on the frame itself
        if (act && present > DH_SLOT) {
            u32 et = rd16(mine, 12), l3 = 14;
            if (et == 0x0081u || et == 0xA888u) {
                l3 = 18;
                if (rd16(mine, 16) == 0x0081u) l3 = 22;
            }
            direct = l3 + (u32)(mine[l3] & 0x0f) * 4 + 8 + 240 + 64 > DH_SLOT;
        }
        if (act) {
            const u32 l0 = len;
            int v = dhcp_one(c, bs, direct ? g : mine, len, frame_dlen(b, l0), frame_now(b, i));
            b.verdict[i] = (u8)v;
            if (len != l0) b.len[i] = len;
        }
        if (direct) nbytes = 0;
        // ---- stage out (every staged frame: a passed frame may have been rewritten, :769) ----
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        if (tile_mode) {
            __syncthreads(); // every thread's writes to the tile are done (and fenced) before the one store reads it
            if (threadIdx.x == 0) {
                asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(b.pkts + (size_t)base * b.stride), "r"(stage_a),
                             "r"(tile_bytes)
                             : "memory");
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); // the tile buffer is reused
            }
        } else if (nbytes) {
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(g), "r"(mine_a), "r"(nbytes) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); // the slot is reused by the next tile
        }
        __syncthreads();
    }
    if (tile_mode && threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    bstats_flush(bs, c.stats);
}

// (A double-buffered variant — two staging slots per thread, the load of tile i+1 issued before the program runs on
// tile i — was measured in round 2 and dropped: 0.97 -> 1.43 ms per 2^22 requests, profiles/r02_notes.md.)

cudaError_t run_dhcp_fastpath(Launcher &L, const DevCtx &c, const DevBatch &b) {
    const int smem = DH_TILE * DH_SLOT;
    if (!L.dhcp_smem_set) { // function attributes are per device: set on the device this context runs on
        cudaError_t e = cudaFuncSetAttribute(k_dhcp_fastpath, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return e;
        L.dhcp_smem_set = 1;
    }
    long want = ((long)b.n + DH_TILE - 1) / DH_TILE;
    long cap = (long)L.num_sms * 4; // 4 x 50 KB of staging per SM
    int grid = (int)(want < cap ? (want < 1 ? 1 : want) : cap);
    prof_begin(L, "k_dhcp_fastpath");
    k_dhcp_fastpath<<<grid, DH_TILE, smem, L.stream>>>(c, b);
    prof_end(L);
    L.launches++;
    return cudaGetLastError();
}

// bng_b200 — session expiry sweep.  The reference declares the NAT timeouts
// (bpf/nat44.c:50-53) but nothing enforces them: its LRU maps evict when full and the Go cleanup loop only
// logs (pkg/nat/manager.go:667-679).  A table that lives in HBM has to be swept: one streaming pass over
// nat_sessions, 64 of every slot's 128 bytes read (key + translation sector, last_seen + state sector).
//
// A session whose idle time exceeds the timeout of its protocol / TCP state goes, together with what hangs
// off it:
//   nat_sessions[key]                               deleted
//   nat_reverse[(dest, nat_ip:nat_port, proto)]     deleted when it still points at this session
//   eim_table[(orig_ip, orig_port, proto)]          ref_count - 1; the mapping goes with its last session
//   subscriber_nat[orig_ip].sessions_active         - 1 (never below 0)
//   nat_stats.sessions_expired                      + 1
//   nat_log_rb                                      one NAT_LOG_SESSION_DELETE record (bpf/nat44.c:72)
// Timeouts: ICMP 60 s, UDP (and anything else) 120 s, TCP ESTABLISHED 7200 s, every other TCP state 240 s.
// Two expiring sessions can share an EIM mapping: the reference count is decremented atomically and the
// thread that takes it to zero erases the mapping.
#include "kernels.h"
#include "progs.cuh"

#define NS 1000000000ull

__device__ __forceinline__ u64 nat_timeout_ns(u32 proto, u32 state) {
    if (proto == 1) return 60ull * NS;
    if (proto == 6) return state == 1 ? 7200ull * NS : 240ull * NS;
    return 120ull * NS;
}

__global__ void __launch_bounds__(256) k_nat_sweep(const __grid_constant__ DevCtx c, u64 now, u32 *n_expired) {
    __shared__ BlockStats bs;
    bstats_init(bs);
    const Tbl &t = c.sessions;
    const u64 slots = (u64)t.mask + 1;
    u32 mine = 0, tombs = 0;
    for (u64 i = blockIdx.x * (u64)blockDim.x + threadIdx.x; i < slots; i += (u64)gridDim.x * blockDim.x) {
        u8 *s = t.slots + i * t.slot_bytes;
        const U256 s0 = ldg256(s);
        const u64 k0 = (u64)s0.w[0] | ((u64)s0.w[1] << 32), k1 = (u64)s0.w[2] | ((u64)s0.w[3] << 32);
        if (k0 >= K_BUSY) {
            tombs += k0 == K_TOMB;
            continue;
        }
        const U256 s1 = ldg256(s + 32); // last_seen 8 | orig_ip 4 | state word 4 | orig_port 2 ...
        const u64 last_seen = (u64)s1.w[0] | ((u64)s1.w[1] << 32);
        const u32 state = s1.w[3] & 0xff, proto = (s1.w[3] >> 8) & 0xff;
        if (now < last_seen || now - last_seen <= nat_timeout_ns(proto, state)) continue;
        if (atomicCAS((u64 *)s, k0, K_TOMB) != k0) continue; // somebody else removed it
        atomicSub(t.count, 1u);
        mine++;
        tombs++;
        const u32 nat_ip = s0.w[4], nat_port = s0.w[5] & 0xffff;
        const u32 orig_ip = s1.w[2], orig_port = s1.w[4] & 0xffff;
        const u32 dest_ip = *(const u32 *)(s + SES_DEST_IP), dest_port = *(const u16 *)(s + SES_DEST_PORT);
        // the reverse entry, when it is still this session's
        u64 rk[2];
        rk[0] = (u64)dest_ip | ((u64)nat_ip << 32);
        rk[1] = (u64)dest_port | ((u64)nat_port << 16) | ((u64)proto << 32);
        u8 *rs = tbl_find<2, true>(c.reverse, rk);
        if (rs && *(volatile const u64 *)(rs + 16) == k0 && *(volatile const u64 *)(rs + 24) == k1 &&
            atomicCAS((u64 *)rs, rk[0], K_TOMB) == rk[0])
            atomicSub(c.reverse.count, 1u);
        // the endpoint mapping loses a reference
        u64 ek = (u64)orig_ip | ((u64)orig_port << 32) | ((u64)proto << 48);
        u8 *m = tbl_find<1, true>(c.eim, &ek);
        if (m) {
            u32 *rc = (u32 *)(m + 32);
            u32 cur = *(volatile u32 *)rc;
            while (cur) {
                const u32 prev = atomicCAS(rc, cur, cur - 1);
                if (prev == cur) break;
                cur = prev;
            }
            if (cur == 1 && atomicCAS((u64 *)m, ek, K_TOMB) == ek) atomicSub(c.eim.count, 1u);
        }
        u64 sk = orig_ip;
        u8 *sub = tbl_find<1, false>(c.sub_nat, &sk);
        u32 sub_id = 0;
        if (sub) {
            sub_id = *(const u32 *)(sub + 32);
            unsigned long long *act = (unsigned long long *)(sub + 40);
            unsigned long long cur = *(volatile unsigned long long *)act;
            while (cur) {
                const unsigned long long prev = atomicCAS(act, cur, cur - 1);
                if (prev == cur) break;
                cur = prev;
            }
        }
        // sweep records carry the marker 0xFFFFFFFE instead of a frame index: the drain orders them by content
        nat_log(c, 0xFFFFFFFEu, now, 2, sub_id, orig_ip, nat_ip, (u16)orig_port, (u16)nat_port, dest_ip, (u16)dest_port, (u8)proto, 0);
    }
    u32 tot = __reduce_add_sync(0xffffffffu, mine), tt = __reduce_add_sync(0xffffffffu, tombs);
    if ((threadIdx.x & 31) == 0) {
        if (tot) {
            bstats_add(bs, ST_NAT_EXPIRED, tot);
            atomicAdd(n_expired, tot);
        }
        if (tt) atomicAdd(n_expired + 1, tt);
    }
    bstats_flush(bs, c.stats);
}

cudaError_t run_nat_sweep(Launcher &L, const DevCtx &c, u64 now, u32 *n_expired) {
    prof_begin(L, "k_nat_sweep");
    k_nat_sweep<<<L.num_sms * 8, 256, 0, L.stream>>>(c, now, n_expired);
    prof_end(L);
    L.launches++;
    return cudaGetLastError();
}

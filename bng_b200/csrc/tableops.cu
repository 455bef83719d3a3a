// bng_b200 — control-plane side of the device hash tables: the kernels behind
// bng_map_update / lookup / delete / dump (the bpf(2) map commands the Go
// managers issue through cilium/ebpf Map.Put/Lookup/Delete).
#include <errno.h>

#include "kernels.h"

template <int KW>
__device__ __forceinline__ void load_key(const Tbl &t, const u8 *kb, u64 *kw) {
    if (t.key_size == 4) {
        kw[0] = *(const u32 *)kb;
    } else {
#pragma unroll
        for (int j = 0; j < KW; j++) kw[j] = *(const u64 *)(kb + 8 * j);
    }
}

__device__ __forceinline__ void copy_bytes(u8 *dst, const u8 *src, u32 n) {
    for (u32 i = 0; i < n; i++) dst[i] = src[i];
}

// value in reference (ABI) layout <-> value as stored in the slot
__device__ __forceinline__ void val_to_slot(const Tbl &t, u8 *slot, const u8 *abi) {
    if (t.vlayout == VL_SESSION) {
        for (u32 i = 0; i < t.value_size; i++) slot[ses_abi_to_slot(i)] = abi[i];
    } else {
        copy_bytes(slot + t.voff, abi, t.value_size);
        // token buckets: rate_bps (value offset 16) is mirrored next to the key, so that the per-frame
        // probe learns "unlimited or not" from the key's own 16 bytes.  Device code never changes the rate.
        if (t.vlayout == VL_QOS) *(u64 *)(slot + QOS_RATE_COPY) = *(const u64 *)(slot + t.voff + 16);
    }
}
__device__ __forceinline__ void val_from_slot(const Tbl &t, u8 *abi, const u8 *slot) {
    if (t.vlayout == VL_SESSION) {
        for (u32 i = 0; i < t.value_size; i++) abi[i] = slot[ses_abi_to_slot(i)];
    } else {
        copy_bytes(abi, slot + t.voff, t.value_size);
    }
}

template <int KW>
__global__ void k_table_op(const __grid_constant__ Tbl t, int op, const u8 *keys, u8 *vals, int *results, u64 n, u32 flags) {
    for (u64 i = blockIdx.x * (u64)blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x) {
        u64 kw[KW];
        load_key<KW>(t, keys + i * t.key_size, kw);
        int r = 0;
        if (kw[0] >= K_BUSY) {
            r = (op == TOP_UPDATE) ? -EINVAL : -ENOENT; // reserved key patterns cannot be stored
        } else if (op == TOP_LOOKUP) {
            const u8 *s = tbl_find<KW, true>(t, kw);
            if (s)
                val_from_slot(t, vals + i * t.value_size, s);
            else
                r = -ENOENT;
        } else if (op == TOP_DELETE) {
            r = tbl_erase<KW>(t, kw) ? 0 : -ENOENT;
        } else {
            const u8 *v = vals + i * t.value_size;
            if (flags == 2) { // BPF_EXIST
                u8 *s = tbl_find<KW, true>(t, kw);
                if (s)
                    val_to_slot(t, s, v);
                else
                    r = -ENOENT;
            } else {
                bool created;
                u8 *s = tbl_find_or_claim<KW>(t, kw, &created);
                if (!s) {
                    r = -E2BIG;
                } else if (!created && flags == 1) { // BPF_NOEXIST
                    r = -EEXIST;
                } else {
                    if (created) // bytes between the key and the value, and the slot tail, stay defined
                        for (u32 z = 8 * KW; z < t.slot_bytes; z += 8) *(u64 *)(s + z) = 0;
                    val_to_slot(t, s, v);
                    if (created) tbl_publish(s, kw[0]);
                }
            }
        }
        results[i] = r;
    }
}

__global__ void k_table_dump(const __grid_constant__ Tbl t, u8 *keys_out, u8 *vals_out, u32 *count_out, u64 cap) {
    u64 slots = (u64)t.mask + 1;
    for (u64 i = blockIdx.x * (u64)blockDim.x + threadIdx.x; i < slots; i += (u64)gridDim.x * blockDim.x) {
        const u8 *s = t.slots + i * t.slot_bytes;
        u64 w0 = *(const u64 *)s;
        if (w0 >= K_BUSY) continue;
        u32 pos = atomicAdd(count_out, 1u);
        if (pos >= cap) continue;
        copy_bytes(keys_out + (u64)pos * t.key_size, s, t.key_size);
        val_from_slot(t, vals_out + (u64)pos * t.value_size, s);
    }
}

cudaError_t run_table_op(Launcher &L, const Tbl &t, int op, const u8 *keys, u8 *vals, int *results, u64 n, u32 flags) {
    if (n == 0) return cudaSuccess;
    int block = 128;
    u64 want = (n + block - 1) / block;
    int grid = (int)(want < (u64)L.num_sms * 8 ? want : (u64)L.num_sms * 8);
    if (t.key_size <= 8)
        k_table_op<1><<<grid, block, 0, L.stream>>>(t, op, keys, vals, results, n, flags);
    else if (t.key_size == 16)
        k_table_op<2><<<grid, block, 0, L.stream>>>(t, op, keys, vals, results, n, flags);
    else
        k_table_op<4><<<grid, block, 0, L.stream>>>(t, op, keys, vals, results, n, flags);
    L.launches++;
    return cudaGetLastError();
}

cudaError_t run_table_dump(Launcher &L, const Tbl &t, u8 *keys_out, u8 *vals_out, u32 *count_out, u64 cap) {
    k_table_dump<<<L.num_sms * 4, 256, 0, L.stream>>>(t, keys_out, vals_out, count_out, cap);
    L.launches++;
    return cudaGetLastError();
}

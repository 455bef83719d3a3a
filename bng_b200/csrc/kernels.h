// bng_b200 — internal interface between the C-ABI layer (ctx.cu) and the
// kernel translation units.
#pragma once
#include "common.cuh"

struct Scratch {
    u32 *key_a, *key_b, *val_a, *val_b; // ordering keys / frame indices, unsorted and grouped
    u32 *qslot;                         // group heads (positions in the grouped arrays)
    void *cub_tmp;
    size_t cub_tmp_bytes;
    u32 *counters; // 16 x u32 of per-run device counters
    u32 cap;       // frames the arrays above can hold
};

// optional per-kernel timing (bng_prof_enable): CUDA events around every launch
struct ProfPending {
    int acc;
    cudaEvent_t a, b;
};

struct Launcher {
    cudaStream_t stream;
    int num_sms;
    Scratch s;
    unsigned long long launches;
    // per-context (= per-device) launch configuration, filled on first use: nothing here may be process-wide,
    // one process can hold contexts on several GPUs
    int dhcp_smem_set;  // cudaFuncAttributeMaxDynamicSharedMemorySize applied on this context's device
    int resolve_bps[6]; // resident blocks per SM of the k_resolve instantiations
    int prof;
    ProfPending pend[32];
    int npend;
    const char *acc_name[32];
    double acc_ms[32];
    unsigned long long acc_n[32];
    int nacc;
};

void prof_begin(Launcher &L, const char *name);
void prof_end(Launcher &L);
void prof_collect(Launcher &L); // call after the stream has been synchronised

size_t sort_temp_bytes(u32 n);

cudaError_t run_antispoof(Launcher &L, const DevCtx &c, const DevBatch &b);
cudaError_t run_qos(Launcher &L, const DevCtx &c, const DevBatch &b, bool egress);
cudaError_t run_nat_egress(Launcher &L, const DevCtx &c, const DevBatch &b);
cudaError_t run_nat_ingress(Launcher &L, const DevCtx &c, const DevBatch &b);
cudaError_t run_nat_hairpin_xdp(Launcher &L, const DevCtx &c, const DevBatch &b);
cudaError_t run_pipeline_up(Launcher &L, const DevCtx &c, const DevBatch &b);
cudaError_t run_pipeline_tc(Launcher &L, const DevCtx &c, const DevBatch &b);
cudaError_t run_dhcp_fastpath(Launcher &L, const DevCtx &c, const DevBatch &b);

// header gather / scatter between a pinned host arena and a compact device copy (hostio.cu)
cudaError_t run_gather_frames(cudaStream_t st, int blocks, const u8 *arena, const u32 *off16, const u32 *len, u32 stride,
                              u32 n, u32 slot, bool tc, u8 *dst, u32 *need);
cudaError_t run_scatter_frames(cudaStream_t st, int blocks, u8 *arena, const u32 *off16, const u32 *need, u32 stride, u32 n,
                               u32 slot, const u8 *src, u32 first_chunk);

// table maintenance (tableops.cu); keys/values/results are device pointers
enum { TOP_UPDATE = 0, TOP_LOOKUP = 1, TOP_DELETE = 2 };
// dir / dir_role: the subscriber directory and which half of it table t feeds (0 none, 1 subscriber_nat, 2 qos_ingress)
cudaError_t run_table_op(Launcher &L, const Tbl &t, int op, const u8 *keys, u8 *vals, int *results, u64 n, u32 flags,
                         const Tbl &dir, int dir_role);
cudaError_t run_dir_clear_half(Launcher &L, const Tbl &dir, int role);
cudaError_t run_epoch_reset(Launcher &L, const Tbl &sessions);
cudaError_t run_table_rebuild(Launcher &L, const Tbl &old_table, const Tbl &empty_table);
// session expiry sweep (sweep.cu); n_expired: device counter, incremented by the number of sessions removed
cudaError_t run_nat_sweep(Launcher &L, const DevCtx &c, u64 now, u32 *n_expired /* [0] expired, [1] tombstones seen */);
cudaError_t run_table_dump(Launcher &L, const Tbl &t, u8 *keys_out, u8 *vals_out, u32 *count_out, u64 cap);

// bng_b200 — frame I/O for BNG_MEM_HOST batches whose arena is pinned host
// memory: instead of copying whole frames over PCIe, gather kernels read only
// the header bytes the programs can touch straight out of the mapped host
// arena (zero-copy, 16 bytes per lane, a frame's chunks on consecutive lanes so
// each frame is one contiguous PCIe read), the programs run on the compact
// device copy, and scatter kernels write the headers back in place.
#include "kernels.h"

// dst[f][0..hb) <- arena[off(f) .. off(f)+min(len, hb)); dlen/dlen0 <- len
__global__ void __launch_bounds__(256) k_gather_frames(const u8 *__restrict__ arena, const u32 *__restrict__ off16,
                                                       const u32 *__restrict__ len, u32 stride, u32 n, u32 hb, u8 *dst,
                                                       u32 *dlen0) {
    const u32 cpf = hb / 16; // chunks per frame
    const u64 total = (u64)n * cpf;
    for (u64 t = blockIdx.x * (u64)blockDim.x + threadIdx.x; t < total; t += (u64)gridDim.x * blockDim.x) {
        u32 f = (u32)(t / cpf), ch = (u32)(t % cpf);
        u32 l = len[f];
        if (ch == 0) dlen0[f] = l;
        if (ch * 16 < l) {
            const u8 *src = arena + (off16 ? (size_t)off16[f] * 16 : (size_t)f * stride) + ch * 16;
            *(uint4 *)(dst + (size_t)f * hb + ch * 16) = *(const uint4 *)src;
        }
    }
}

// arena[off(f) .. ) <- dst[f][0..min(len0, hb)) for frames the program may have written
__global__ void __launch_bounds__(256) k_scatter_frames(u8 *__restrict__ arena, const u32 *__restrict__ off16,
                                                        const u32 *__restrict__ dlen0, u32 stride, u32 n, u32 hb,
                                                        const u8 *__restrict__ src, u32 first_chunk) {
    const u32 cpf = hb / 16;
    const u64 total = (u64)n * cpf;
    for (u64 t = blockIdx.x * (u64)blockDim.x + threadIdx.x; t < total; t += (u64)gridDim.x * blockDim.x) {
        u32 f = (u32)(t / cpf), ch = (u32)(t % cpf);
        if (ch < first_chunk) continue;
        u32 l = dlen0[f];
        if (ch * 16 < l) {
            u8 *d = arena + (off16 ? (size_t)off16[f] * 16 : (size_t)f * stride) + ch * 16;
            *(uint4 *)d = *(const uint4 *)(src + (size_t)f * hb + ch * 16);
        }
    }
}

cudaError_t run_gather_frames(cudaStream_t st, int num_sms, const u8 *arena, const u32 *off16, const u32 *len, u32 stride,
                              u32 n, u32 hb, u8 *dst, u32 *dlen0) {
    k_gather_frames<<<num_sms * 8, 256, 0, st>>>(arena, off16, len, stride, n, hb, dst, dlen0);
    return cudaGetLastError();
}

cudaError_t run_scatter_frames(cudaStream_t st, int num_sms, u8 *arena, const u32 *off16, const u32 *dlen0, u32 stride, u32 n,
                               u32 hb, const u8 *src, u32 first_chunk) {
    k_scatter_frames<<<num_sms * 8, 256, 0, st>>>(arena, off16, dlen0, stride, n, hb, src, first_chunk);
    return cudaGetLastError();
}

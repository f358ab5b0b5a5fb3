// tcgen05 / TMA / mbarrier PTX wrappers and UMMA shared-memory descriptors shared by the
// tensor-core kernels (conv_tc.cu, stem_tc.cu).  sm_100a only.
#pragma once
#include "common.cuh"
#include <cudaTypedefs.h>
#include <cuda_bf16.h>

namespace {

// ---------------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// Bounded wait: a protocol bug traps (-> CUDA error on the host) instead of hanging the GPU.  try_wait suspends the
// thread in hardware for up to the time hint, so a waiting warp costs few issue slots (ncu on the stem kernels: the
// former clock64()-bounded spin was ~180 of 1000 instructions per tile); the bound counts attempts instead.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    for (uint32_t spin = 0; spin < (1u << 20); ++spin) {
        asm volatile(
            "{\n\t"
            ".reg .pred P1;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2, %3;\n\t"
            "selp.u32 %0, 1, 0, P1;\n\t"
            "}" : "=r"(done) : "r"(bar), "r"(parity), "r"(20000u) : "memory");
        if (done) return;
    }
    printf("dpc_b200: mbarrier wait timed out (block %d,%d,%d thread %d bar 0x%x parity %u)\n", blockIdx.x,
           blockIdx.y, blockIdx.z, threadIdx.x, bar, parity);
    __trap();
}
__device__ __forceinline__ void tma_load_5d(const CUtensorMap* map, uint32_t dst, uint32_t bar, int c0, int c1,
                                            int c2, int c3, int c4) {
    asm volatile(
        "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4) : "memory");
}
__device__ __forceinline__ void tma_load_2d(const CUtensorMap* map, uint32_t dst, uint32_t bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// asynchronous: the registers are defined only after tmem_ld_wait()
__device__ __forceinline__ void tmem_ld32_nowait(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
          "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    tmem_ld32_nowait(taddr, v);
    tmem_ld_wait();
}
__device__ __forceinline__ bool elect_one() {
    uint32_t pred = 0;
    asm volatile(
        "{\n\t"
        ".reg .pred P;\n\t"
        "elect.sync _|P, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, P;\n\t"
        "}" : "=r"(pred));
    return pred != 0;
}

// Row-per-lane epilogue stores with full sectors.  After tcgen05.ld every lane holds ONE output row; a plain STG.128
// per lane touches 32 different rows, i.e. 32 half-filled 32-byte sectors per instruction (measured on the stem
// forward: 2x the L2 write sectors, L2 72 % busy).  Here `o` holds NG groups of 8 consecutive columns (two float4)
// of this lane's row; lanes 2k / 2k+1 exchange one float4 per group and store
// [row 2k | row 2k+1] x [cols 0-3 | cols 4-7], so each instruction of the pair fills whole sectors of one row.
// `o` must be computed by ALL lanes (shuffles); `accumulate` adds the existing values first.
template <int NG>
__device__ __forceinline__ void pair_store_rows(const float4 (&o)[2 * NG], float* __restrict__ y, long long row, bool valid,
                                                int lane, int ld, int c0, bool accumulate) {
    const long long row_p = __shfl_xor_sync(0xffffffffu, row, 1);
    const bool valid_p = __shfl_xor_sync(0xffffffffu, (int)valid, 1) != 0;
    const bool odd = (lane & 1) != 0;
    const long long row_e = odd ? row_p : row, row_o = odd ? row : row_p;
    const bool val_e = odd ? valid_p : valid, val_o = odd ? valid : valid_p;
    float* pe = y + (val_e ? row_e : 0) * ld + c0 + (odd ? 4 : 0);
    float* po = y + (val_o ? row_o : 0) * ld + c0 + (odd ? 4 : 0);
#pragma unroll
    for (int g = 0; g < NG; ++g) {
        const float4 a = o[2 * g], b = o[2 * g + 1];
        const float4 send = odd ? a : b;
        float4 recv;
        recv.x = __shfl_xor_sync(0xffffffffu, send.x, 1);
        recv.y = __shfl_xor_sync(0xffffffffu, send.y, 1);
        recv.z = __shfl_xor_sync(0xffffffffu, send.z, 1);
        recv.w = __shfl_xor_sync(0xffffffffu, send.w, 1);
        float4 s1 = odd ? recv : a, s2 = odd ? b : recv;
        if (val_e) {
            float4* d = reinterpret_cast<float4*>(pe + 8 * g);
            if (accumulate) { const float4 c = *d; s1.x += c.x; s1.y += c.y; s1.z += c.z; s1.w += c.w; }
            *d = s1;
        }
        if (val_o) {
            float4* d = reinterpret_cast<float4*>(po + 8 * g);
            if (accumulate) { const float4 c = *d; s2.x += c.x; s2.y += c.y; s2.z += c.z; s2.w += c.w; }
            *d = s2;
        }
    }
}

// K-major, 128B-swizzled operand tile: rows at 128-byte pitch, 8-row atoms 1024 bytes apart.
__device__ __forceinline__ uint64_t make_kmajor_sw128_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);        // start address
    d |= (uint64_t)1 << 16;                        // leading byte offset (unused for swizzled K-major)
    d |= (uint64_t)(1024 >> 4) << 32;              // stride byte offset: 8 rows * 128 B
    d |= (uint64_t)1 << 46;                        // descriptor version (Blackwell)
    d |= (uint64_t)2 << 61;                        // SWIZZLE_128B
    return d;
}
// MN-major, 128B-swizzled: each smem row is one K index holding 64 contiguous MN elements (128 B);
// 8 K-rows form a swizzle atom (SBO = 1024 B); 64-element MN groups are `lbo_bytes` apart.
__device__ __forceinline__ uint64_t make_mnmajor_sw128_desc(uint32_t saddr, uint32_t lbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}

// K-major, NO swizzle ("interleaved" 8x16-byte core matrices): core matrices adjacent in K are
// `lbo_bytes` apart, 8-row groups adjacent in M/N are `sbo_bytes` apart.  For tiles written by
// CUDA cores (the stem's im2col tile).  The same physical tile read as an MN-major operand
// (MN = the former K, K = the former rows) swaps the two offsets.
__device__ __forceinline__ uint64_t make_noswizzle_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    return d;                                      // layout type 0 = SWIZZLE_NONE
}

}  // namespace

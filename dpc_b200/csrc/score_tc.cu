// The score matmul of DPC_RNN.forward:  score[M, N] = pred[M, K] . feature_inf[N, K]^T  (K = feature size 256, M = N =
// B * pred_step * last_size^2 = 6144 at BASELINE config 2), fp32 output, split-precision operands (3 MMA passes).
//
// The generic conv/GEMM kernel (conv_tc.cu) ran this as 1152 one-tile CTAs: with K = 256 a tile is only 4 k-blocks, its
// 128 x 256 accumulators need all 512 TMEM columns (so the two co-resident CTAs serialise), every tile re-fetches its A
// rows, and the 151 MB fp32 output -- the actual bound, ~25 us at HBM speed -- is written by 4 epilogue warps per CTA
// with nothing overlapping it: 111 us = 18 % of that bound (round-1 VERDICT item 8).  Here:
//   * one persistent CTA per (128-row M tile, range of N tiles): the M tile's operand planes (K = 256: 4 x [hi | lo] x
//     16 KB = 128 KB) stay RESIDENT in shared memory for the whole sweep, only B streams (3-stage ring of 32 KB
//     [B_hi ; B_lo] k-blocks): L2 -> SM traffic per output tile drops from 384 KB to 128 KB;
//   * 128 x 128 tiles with double-buffered TMEM accumulators (main | cross-term, 2 x 256 columns): the 8 epilogue warps
//     drain tile i (tcgen05.ld -> add -> full-sector stores) while the MMAs of tile i + 1 run;
//   * A_hi x [B_hi ; B_lo] is ONE N = 256 instruction (the two B planes are adjacent in a stage, the two accumulators
//     adjacent in TMEM), A_lo x B_hi one N = 128 instruction.
// Replaces torch.matmul at dpc/model_3d.py:83.
#include "tc_common.cuh"

namespace {

constexpr int SC_K = 256, SC_KB = SC_K / 64;         // feature size; 64-column k-blocks
constexpr int SC_BN = 128;
constexpr int SC_STAGES = 3;
constexpr uint32_t SC_TILE = 128 * 128;              // one [128 rows x 64 bf16/fp16] plane tile: 16 KB
constexpr int SC_THREADS = 64 + 32 * 8;              // producer + MMA warp + 8 epilogue warps

struct ScMaps { CUtensorMap a_hi, a_lo, b_hi, b_lo; };

struct ScParams {
    int M, N;
    int m_tiles, n_tiles, n_splits, n_per_split;
    int f16;                                         // operand planes: fp16 pairs (1) or bf16 pairs (0)
};

__device__ __forceinline__ void sc_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

__global__ void __launch_bounds__(SC_THREADS, 1)
score_gemm_kernel(const __grid_constant__ ScMaps maps, const ScParams p, float* __restrict__ C) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // smem: [A resident: 4 k-blocks x (hi | lo)] [ring: 3 x (B_hi | B_lo)] [barriers]
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t ring = base + SC_KB * 2u * SC_TILE;
    const uint32_t bar_base = ring + SC_STAGES * 2u * SC_TILE;
    auto full = [&](int s) { return bar_base + 8u * s; };
    auto empty = [&](int s) { return bar_base + 8u * (SC_STAGES + s); };
    const uint32_t a_full = bar_base + 8u * (2 * SC_STAGES);
    auto tm_full = [&](int b) { return bar_base + 8u * (2 * SC_STAGES + 1 + b); };
    auto tm_empty = [&](int b) { return bar_base + 8u * (2 * SC_STAGES + 3 + b); };
    const uint32_t tmem_ptr_addr = bar_base + 8u * (2 * SC_STAGES + 5);
    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.a_hi) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.a_lo) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.b_hi) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.b_lo) : "memory");
        for (int s = 0; s < SC_STAGES; ++s) { mbar_init(full(s), 1); mbar_init(empty(s), 1); }
        mbar_init(a_full, 1);
        for (int b = 0; b < 2; ++b) { mbar_init(tm_full(b), 1); mbar_init(tm_empty(b), 8); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) tmem_alloc(tmem_ptr_addr, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<const uint32_t*>(smem_raw + (tmem_ptr_addr - smem_u32(smem_raw)));
    const int mt = (int)blockIdx.x / p.n_splits, ns = (int)blockIdx.x - mt * p.n_splits;
    const int nt0 = ns * p.n_per_split;
    const int nt1 = nt0 + p.n_per_split < p.n_tiles ? nt0 + p.n_per_split : p.n_tiles;
    const int my_tiles = (mt < p.m_tiles && nt1 > nt0) ? nt1 - nt0 : 0;
    const int m0 = mt * 128;

    if (warp == 0) {
        if (elect_one() && my_tiles > 0) {
            mbar_expect_tx(a_full, SC_KB * 2u * SC_TILE);
            for (int kb = 0; kb < SC_KB; ++kb) {
                tma_load_2d(&maps.a_hi, base + kb * 2u * SC_TILE, a_full, kb * 64, m0);
                tma_load_2d(&maps.a_lo, base + kb * 2u * SC_TILE + SC_TILE, a_full, kb * 64, m0);
            }
            int s = 0; uint32_t ph = 0;
            for (int i = 0; i < my_tiles; ++i) {
                const int n0 = (nt0 + i) * SC_BN;
                for (int kb = 0; kb < SC_KB; ++kb) {
                    mbar_wait(empty(s), ph ^ 1u);
                    mbar_expect_tx(full(s), 2u * SC_TILE);
                    const uint32_t sb = ring + s * 2u * SC_TILE;
                    tma_load_2d(&maps.b_hi, sb, full(s), kb * 64, n0);
                    tma_load_2d(&maps.b_lo, sb + SC_TILE, full(s), kb * 64, n0);
                    if (++s == SC_STAGES) { s = 0; ph ^= 1u; }
                }
            }
        }
    } else if (warp == 1) {
        if (elect_one() && my_tiles > 0) {
            // D = f32; A = B = bf16 (bits 7, 10) or fp16 (0); K-major; M = 128; N = 256 / 128
            const uint32_t fmt = (1u << 4) | (p.f16 ? 0u : ((1u << 7) | (1u << 10))) | ((128u >> 4) << 24);
            const uint32_t idesc2 = fmt | ((uint32_t)(256 >> 3) << 17), idesc1 = fmt | ((uint32_t)(128 >> 3) << 17);
            mbar_wait(a_full, 0);
            int s = 0; uint32_t ph = 0;
            for (int i = 0; i < my_tiles; ++i) {
                const int buf = i & 1;
                const uint32_t td = tmem_base + (uint32_t)(buf * 256), tcx = td + 128u;
                mbar_wait(tm_empty(buf), (((uint32_t)i >> 1) & 1u) ^ 1u);
                tc_fence_after();
#pragma unroll
                for (int kb = 0; kb < SC_KB; ++kb) {
                    mbar_wait(full(s), ph);
                    tc_fence_after();
                    const uint32_t sa = base + kb * 2u * SC_TILE, sb = ring + s * 2u * SC_TILE;
                    const uint64_t ahi = make_kmajor_sw128_desc(sa), alo = make_kmajor_sw128_desc(sa + SC_TILE);
                    const uint64_t bhi = make_kmajor_sw128_desc(sb);
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const uint64_t ko = (uint64_t)(k * 2);
                        umma_bf16(td, ahi + ko, bhi + ko, idesc2, (kb | k) ? 1u : 0u);      // A_hi x [B_hi ; B_lo] -> main | cross
                        umma_bf16(tcx, alo + ko, bhi + ko, idesc1, 1u);                     // A_lo x B_hi -> cross
                    }
                    umma_commit(empty(s));
                    if (++s == SC_STAGES) { s = 0; ph ^= 1u; }
                }
                umma_commit(tm_full(buf));
            }
        }
    } else {
        // epilogue: 8 warps = 4 TMEM lane quarters x two 64-column halves
        const int q = warp & 3, c_half = ((warp - 2) >> 2) * 64;
        const long long row = (long long)m0 + q * 32 + lane;
        const bool valid = row < p.M;
        for (int i = 0; i < my_tiles; ++i) {
            const int buf = i & 1;
            const uint32_t td = tmem_base + (uint32_t)(buf * 256), tcx = td + 128u;
            const int n0 = (nt0 + i) * SC_BN + c_half;
            mbar_wait(tm_full(buf), ((uint32_t)i >> 1) & 1u);
            tc_fence_after();
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                uint32_t v[32], u[32];
                tmem_ld32_nowait(td + ((uint32_t)(q * 32) << 16) + (uint32_t)(c_half + 32 * c), v);
                tmem_ld32_nowait(tcx + ((uint32_t)(q * 32) << 16) + (uint32_t)(c_half + 32 * c), u);
                tmem_ld_wait();
                if (c == 1) {                                         // both chunks are in registers: hand the buffer back
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) sc_arrive(tm_empty(buf));
                }
                const int nc = n0 + 32 * c;
                if (nc + 32 <= p.N && (p.N & 3) == 0) {
                    float4 o[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        o[j] = make_float4(__uint_as_float(v[4 * j]) + __uint_as_float(u[4 * j]),
                                           __uint_as_float(v[4 * j + 1]) + __uint_as_float(u[4 * j + 1]),
                                           __uint_as_float(v[4 * j + 2]) + __uint_as_float(u[4 * j + 2]),
                                           __uint_as_float(v[4 * j + 3]) + __uint_as_float(u[4 * j + 3]));
                    pair_store_rows<4>(o, C, row, valid, lane, p.N, nc, false);          // full 32-byte sectors
                } else if (valid) {
#pragma unroll
                    for (int j = 0; j < 32; ++j)
                        if (nc + j < p.N) C[row * p.N + nc + j] = __uint_as_float(v[j]) + __uint_as_float(u[j]);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, 512);
}

PFN_cuTensorMapEncodeTiled_v12000 sc_encode() {
    static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
    }
    return fn;
}

}  // namespace

// score[M, N] = A[M, 256] . B[N, 256]^T (fp32, row-major, ld = N) from split planes; f16: fp16 pairs (else bf16 pairs).
// The persistent A-resident schedule; K must be 256 (the feature size of r18 / r34) -- other K: dpc_gemm_nt_split_tc.
extern "C" int dpc_score_matmul_tc(int M, int N, int K, const void* a_hi, const void* a_lo, const void* b_hi, const void* b_lo,
                                   int f16, float* C, void* stream) {
    DPC_REQUIRE(M > 0 && N > 0 && K == SC_K, "dpc_score_matmul_tc: K must be %d (got M %d N %d K %d)", SC_K, M, N, K);
    DPC_REQUIRE(a_hi && a_lo && b_hi && b_lo && C, "dpc_score_matmul_tc: null pointer");
    auto enc = sc_encode();
    DPC_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled entry point not available");
    ScMaps maps;
    const cuuint32_t es[2] = {1, 1}, box[2] = {64, 128};
    const cuuint64_t gs[1] = {(cuuint64_t)K * 2};
    const void* ptrs[4] = {a_hi, a_lo, b_hi, b_lo};
    CUtensorMap* ms[4] = {&maps.a_hi, &maps.a_lo, &maps.b_hi, &maps.b_lo};
    for (int i = 0; i < 4; ++i) {
        const cuuint64_t gd[2] = {(cuuint64_t)K, (cuuint64_t)(i < 2 ? M : N)};
        CUresult r = enc(ms[i], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptrs[i]), gd, gs, box, es,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        DPC_REQUIRE(r == CUDA_SUCCESS, "dpc_score_matmul_tc: cuTensorMapEncodeTiled failed (%d)", (int)r);
    }
    ScParams p;
    p.M = M; p.N = N; p.f16 = f16 ? 1 : 0;
    p.m_tiles = (M + 127) / 128;
    p.n_tiles = (N + SC_BN - 1) / SC_BN;
    const int sms = dpc_num_sms();
    p.n_splits = sms / p.m_tiles;
    if (p.n_splits < 1) p.n_splits = 1;
    if (p.n_splits > p.n_tiles) p.n_splits = p.n_tiles;
    p.n_per_split = (p.n_tiles + p.n_splits - 1) / p.n_splits;
    p.n_splits = (p.n_tiles + p.n_per_split - 1) / p.n_per_split;
    const size_t smem = 1024 + (size_t)SC_KB * 2 * SC_TILE + (size_t)SC_STAGES * 2 * SC_TILE + 8 * (2 * SC_STAGES + 6);
    DPC_CUDA(cudaFuncSetAttribute(score_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    score_gemm_kernel<<<p.m_tiles * p.n_splits, SC_THREADS, smem, as_stream(stream)>>>(maps, p, C);
    DPC_LAUNCH_CHECK();
    return DPC_OK;
}

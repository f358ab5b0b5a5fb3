// Stem convolution (conv1: 3 -> 64, 1x7x7, stride (1,2,2), pad (0,3,3)) as a STRIDE-1 4x4 convolution over the
// 2x2 space-to-depth input, on the tcgen05 tensor cores.  sm_100a only.
//
//   y[ho,wo] = sum_{c,kh,kw} w[c,kh,kw] x[c, 2ho+kh-3, 2wo+kw-3]
//            = sum_{a,b in 0..3} sum_{c,r,s} W2[a,b][c,r,s] X2[ho+a-2, wo+b-2][c,r,s]
//   X2[h2,w2][(c,r,s)] = x[c, 2h2+r, 2w2+s],   W2[a,b][c,r,s] = w[c, 2a+r-1, 2b+s-1]  (0 where an index is -1)
//
// X2 is stored channels-last as split-bf16 planes with 16 channels per pixel (12 real + 4 zero): one pixel = one
// 32-byte row = exactly one UMMA k-step (K = 16).  This file holds the shared pieces: the pack kernels (input planes,
// filter bank) and the stand-alone wgrad kernel (used when the fused backward of stem_pool.cu does not fit shared memory);
// the forward and the recomputing backward live in stem_pool.cu.
//
// Replaces nn.Conv3d(3, 64, (1,7,7), stride (1,2,2), padding (0,3,3)) at backbone/resnet_2d3d.py:211.
#include "tc_common.cuh"

namespace {

constexpr int S2D_CH = 16, S2D_TAPS = 16, S2D_BN = 64;
constexpr uint32_t S2D_W_BYTES = S2D_TAPS * 2 * S2D_BN * 32;      // 64 KB: per tap [W_hi (64 rows) ; W_lo (64 rows)] x 32 B

__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

// x [NB,3,T,H,W] fp32 -> X2 planes [NB,T,H/2,W/2,16] bf16 (hi = bf16(v), lo = bf16(v - hi))
__global__ void stem_s2d_pack_kernel(const float* __restrict__ x, uint4* __restrict__ hi, uint4* __restrict__ lo, int NB,
                                     int T, int H, int W) {
    const int H2 = H >> 1, W2 = W >> 1;
    const long long total = (long long)NB * T * H2 * W2;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int w2 = (int)(i % W2);
        const int h2 = (int)((i / W2) % H2);
        const int t = (int)((i / ((long long)W2 * H2)) % T);
        const int n = (int)(i / ((long long)W2 * H2 * T));
        uint32_t ph[8], pl[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) { ph[j] = 0u; pl[j] = 0u; }
#pragma unroll
        for (int c = 0; c < 3; ++c)
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                const float2 v = *reinterpret_cast<const float2*>(x + ((((size_t)n * 3 + c) * T + t) * H + 2 * h2 + r) * W + 2 * w2);
                const __nv_bfloat16 h0 = __float2bfloat16_rn(v.x), h1 = __float2bfloat16_rn(v.y);
                const __nv_bfloat16 l0 = __float2bfloat16_rn(v.x - __bfloat162float(h0));
                const __nv_bfloat16 l1 = __float2bfloat16_rn(v.y - __bfloat162float(h1));
                ph[c * 2 + r] = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
                pl[c * 2 + r] = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
            }
        hi[2 * i] = make_uint4(ph[0], ph[1], ph[2], ph[3]);
        hi[2 * i + 1] = make_uint4(ph[4], ph[5], ph[6], ph[7]);
        lo[2 * i] = make_uint4(pl[0], pl[1], pl[2], pl[3]);
        lo[2 * i + 1] = make_uint4(pl[4], pl[5], pl[6], pl[7]);
    }
}

// w [64,3,1,7,7] fp32 -> wp [16 taps][hi | lo][64 co][16 ch] bf16
__global__ void stem_s2d_wpack_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ wp) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= S2D_TAPS * S2D_BN * S2D_CH) return;
    const int ch = i % S2D_CH, co = (i / S2D_CH) % S2D_BN, tap = i / (S2D_CH * S2D_BN);
    const int a = tap >> 2, b = tap & 3;
    float v = 0.f;
    if (ch < 12) {
        const int c = ch >> 2, r = (ch >> 1) & 1, s = ch & 1;
        const int kh = 2 * a + r - 1, kw = 2 * b + s - 1;
        if (kh >= 0 && kw >= 0) v = w[(co * 3 + c) * 49 + kh * 7 + kw];
    }
    const __nv_bfloat16 h = __float2bfloat16_rn(v);
    const __nv_bfloat16 l = __float2bfloat16_rn(v - __bfloat162float(h));
    wp[((size_t)(tap * 2 + 0) * S2D_BN + co) * S2D_CH + ch] = h;
    wp[((size_t)(tap * 2 + 1) * S2D_BN + co) * S2D_CH + ch] = l;
}

// K-major SWIZZLE_32B descriptors: rows of 32 bytes, 8-row atoms 256 bytes apart.  High word = SBO (256 >> 4) |
// version 1 (bit 46) | layout type 6 (SWIZZLE_32B, bits 61-63); low word = start address >> 4 | LBO field (unused).
constexpr uint64_t S2D_DHI = 0xC0004010ull << 32;

// =============================================================================================
// wgrad of conv1 from the same planes:  dW2[co][(a,b)][ch] = sum_pos dY[pos, co] * X2[pos + (a-2, b-2)][ch]
//
// K = positions.  Tiles as in the forward; the dY rows of a tile come in one SWIZZLE_128B box whose out-of-frame
// columns / rows are TMA zero fill, the X2 patch in one SWIZZLE_32B box.  Both operands are MN-major views:
//   A (M = 128) = [dY_hi^T ; dY_lo^T]: the operand's two 64-channel groups are the two planes (LBO = plane pitch);
//   B (N = 64)  = X2^T of the four taps (a, b = 0..3): four 16-channel groups one patch row (32 B) apart.
// One instruction therefore yields dY_hi*X and dY_lo*X for four taps; issuing it for X_hi and X_lo gives all four
// split products (rows 0-63: hi*hi + hi*lo, rows 64-127: lo*hi + lo*lo), which the drain adds.  4 accumulators of
// 64 columns (one per a), flushed to dw with atomics every `chain` tiles.
// =============================================================================================
struct S2dWgMaps { CUtensorMap x_hi, x_lo, y_hi, y_lo; };

struct S2dWgParams {
    int PW, bhr_x, bhr_y, Ho, Wo, T;
    int tiles_per_frame, total_tiles;
    int xpatch_bytes, ypatch_bytes;
    int chain;
};

__global__ void __launch_bounds__(192, 1)
stem_s2d_wgrad_kernel(const __grid_constant__ S2dWgMaps maps, const S2dWgParams hp, float* __restrict__ dw) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // smem: 2 x [dy_hi | dy_lo | x_hi | x_lo] [barriers]
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bufsz = 2u * (uint32_t)hp.ypatch_bytes + 2u * (uint32_t)hp.xpatch_bytes;
    const uint32_t bar_base = base + 2u * bufsz;
    auto full = [&](int b) { return bar_base + 8u * b; };
    auto empty = [&](int b) { return bar_base + 8u * (2 + b); };
    const uint32_t acc_full = bar_base + 32u, acc_empty = bar_base + 40u, tmem_ptr_addr = bar_base + 48u;
    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.x_hi) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.x_lo) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.y_hi) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.y_lo) : "memory");
        for (int b = 0; b < 2; ++b) { mbar_init(full(b), 1); mbar_init(empty(b), 1); }
        mbar_init(acc_full, 1);
        mbar_init(acc_empty, 4);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) tmem_alloc(tmem_ptr_addr, 256);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<const uint32_t*>(smem_raw + (tmem_ptr_addr - smem_u32(smem_raw)));
    const int my_tiles = ((int)blockIdx.x < hp.total_tiles) ? (hp.total_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
    const int chains = (my_tiles + hp.chain - 1) / hp.chain;

    auto tile_origin = [&](int i, int& n, int& t, int& f0, int& hrow0) {
        const int tile = (int)blockIdx.x + i * (int)gridDim.x;
        const int frame = tile / hp.tiles_per_frame;
        f0 = (tile - frame * hp.tiles_per_frame) * 128;
        hrow0 = f0 / hp.PW;
        n = frame / hp.T; t = frame - n * hp.T;
    };

    if (warp == 0) {
        if (elect_one()) {
            const uint32_t tx = 2u * (uint32_t)(hp.bhr_x * hp.PW) * 32u + 2u * (uint32_t)(hp.bhr_y * hp.PW) * 128u;
            for (int i = 0; i < my_tiles; ++i) {
                const int b = i & 1;
                int n, t, f0, hrow0;
                tile_origin(i, n, t, f0, hrow0);
                mbar_wait(empty(b), (((uint32_t)i >> 1) & 1u) ^ 1u);
                mbar_expect_tx(full(b), tx);
                const uint32_t sy = base + b * bufsz, sx = sy + 2u * hp.ypatch_bytes;
                tma_load_5d(&maps.y_hi, sy, full(b), 0, 0, hrow0, t, n);
                tma_load_5d(&maps.y_lo, sy + hp.ypatch_bytes, full(b), 0, 0, hrow0, t, n);
                tma_load_5d(&maps.x_hi, sx, full(b), 0, -2, hrow0 - 2, t, n);
                tma_load_5d(&maps.x_lo, sx + hp.xpatch_bytes, full(b), 0, -2, hrow0 - 2, t, n);
            }
        }
    } else if (warp == 1) {
        if (elect_one()) {
            // D = f32, A = B = bf16, both MN-major (bits 15, 16), N = 64, M = 128
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((64u >> 3) << 17) | ((128u >> 4) << 24);
            constexpr uint64_t DHI128 = 0x40004040ull << 32;   // SWIZZLE_128B, SBO = 1024 B
            const uint32_t ylbo = ((uint32_t)hp.ypatch_bytes >> 4) << 16;      // A: dY_lo group = dY_hi group + one plane
            const uint32_t xlbo = 2u << 16;                                    // B: next tap = next patch row (32 B)
            const uint32_t xpatch16 = (uint32_t)hp.xpatch_bytes >> 4;
            const uint32_t arow16 = (uint32_t)hp.PW * 2u;                      // one tap row a -> PW patch rows of 32 B
            for (int i = 0; i < my_tiles; ++i) {
                const int b = i & 1;
                const int in_chain = i % hp.chain;
                int n, t, f0, hrow0;
                tile_origin(i, n, t, f0, hrow0);
                const uint32_t rowoff = (uint32_t)(f0 - hrow0 * hp.PW);
                const uint32_t y16 = (((base + b * bufsz) >> 4) + rowoff * 8u) | ylbo;
                const uint32_t x16 = (((base + b * bufsz + 2u * hp.ypatch_bytes) >> 4) + rowoff * 2u) | xlbo;
                if (in_chain == 0 && i > 0) mbar_wait(acc_empty, (((uint32_t)(i / hp.chain) - 1u) & 1u));
                mbar_wait(full(b), ((uint32_t)i >> 1) & 1u);
                tc_fence_after();
#pragma unroll
                for (int k = 0; k < 8; ++k) {                  // UMMA_K = 16 positions: 2048 B of dY rows, 512 B of X2 rows
                    const uint64_t ya = DHI128 | (uint64_t)(y16 + k * 128);
#pragma unroll
                    for (int a = 0; a < 4; ++a) {
                        const uint32_t xa = x16 + (uint32_t)a * arow16 + k * 32;
                        const uint32_t d = tmem_base + (uint32_t)(a * 64);
                        umma_bf16(d, ya, S2D_DHI | (uint64_t)xa, idesc, (in_chain | k) ? 1u : 0u);
                        umma_bf16(d, ya, S2D_DHI | (uint64_t)(xa + xpatch16), idesc, 1u);
                    }
                }
                umma_commit(empty(b));
                if (in_chain == hp.chain - 1 || i == my_tiles - 1) umma_commit(acc_full);
            }
        }
    } else {
        // drain: TMEM lane = (plane of dY, co); column = (b, ch) of tap row a
        const int q = warp & 3;
        const int co = (q * 32 + lane) & 63;
        float* dwc = dw + (size_t)co * 147;
        for (int c = 0; c < chains; ++c) {
            mbar_wait(acc_full, (uint32_t)c & 1u);
            tc_fence_after();
#pragma unroll
            for (int a = 0; a < 4; ++a) {
                uint32_t v[32], u[32];
                tmem_ld32_nowait(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(a * 64), v);
                tmem_ld32_nowait(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(a * 64 + 32), u);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 64; ++j) {
                    const int bb = j >> 4, ch = j & 15;
                    const int cc = ch >> 2, r = (ch >> 1) & 1, s = ch & 1;
                    const int kh = 2 * a + r - 1, kw = 2 * bb + s - 1;
                    if (ch < 12 && kh >= 0 && kw >= 0)
                        atomicAdd(dwc + cc * 49 + kh * 7 + kw, __uint_as_float(j < 32 ? v[j] : u[j - 32]));
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(acc_empty);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, 256);
}

PFN_cuTensorMapEncodeTiled_v12000 s2d_encode() {
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

// x [NB,3,T,H,W] fp32 -> split-bf16 space-to-depth planes x2_hi / x2_lo [NB,T,H/2,W/2,16]  (H, W even)
extern "C" int dpc_stem_s2d_pack(const float* x, void* x2_hi, void* x2_lo, int NB, int T, int H, int W, void* stream) {
    DPC_REQUIRE(x && x2_hi && x2_lo && NB > 0 && T > 0 && H > 0 && W > 0, "dpc_stem_s2d_pack: bad args");
    DPC_REQUIRE((H & 1) == 0 && (W & 1) == 0, "dpc_stem_s2d_pack: H (%d) and W (%d) must be even", H, W);
    const long long total = (long long)NB * T * (H / 2) * (W / 2);
    const long long blocks = (total + 255) / 256, cap = (long long)dpc_num_sms() * 16;
    stem_s2d_pack_kernel<<<(int)(blocks < cap ? blocks : cap), 256, 0, as_stream(stream)>>>(
        x, reinterpret_cast<uint4*>(x2_hi), reinterpret_cast<uint4*>(x2_lo), NB, T, H, W);
    DPC_LAUNCH_CHECK();
    return DPC_OK;
}

// w [64,3,1,7,7] fp32 -> the packed split-bf16 filter bank `wp` (32768 bf16: 16 taps x [W_hi ; W_lo] x 64 x 16) that the
// pooled-stem kernels (stem_pool.cu) keep resident in shared memory
extern "C" int dpc_stem_s2d_wpack(const float* w, void* wp, void* stream) {
    DPC_REQUIRE(w && wp, "dpc_stem_s2d_wpack: bad args");
    stem_s2d_wpack_kernel<<<(S2D_TAPS * S2D_BN * S2D_CH + 255) / 256, 256, 0, as_stream(stream)>>>(w, reinterpret_cast<__nv_bfloat16*>(wp));
    DPC_LAUNCH_CHECK();
    return DPC_OK;
}

// dw [64,3,1,7,7] = wgrad of conv1 from the space-to-depth planes and the split-bf16 planes of dy [NB,T,H/2,W/2,64]
extern "C" int dpc_stem_conv_wgrad_s2d(const void* x2_hi, const void* x2_lo, const void* dy_hi, const void* dy_lo, float* dw,
                                       int NB, int T, int H, int W, void* stream) {
    DPC_REQUIRE(x2_hi && x2_lo && dy_hi && dy_lo && dw && NB > 0 && T > 0 && H > 0 && W > 0, "dpc_stem_conv_wgrad_s2d: bad args");
    DPC_REQUIRE((H & 1) == 0 && (W & 1) == 0, "dpc_stem_conv_wgrad_s2d: H (%d) and W (%d) must be even", H, W);
    cudaStream_t st = as_stream(stream);
    const int Ho = H / 2, Wo = W / 2;
    S2dWgParams hp;
    memset(&hp, 0, sizeof(hp));
    hp.PW = Wo + 3;
    hp.bhr_x = 4 + (130 + hp.PW - 1) / hp.PW;
    hp.bhr_y = (127 + 2 * hp.PW - 1) / hp.PW;
    DPC_REQUIRE(hp.PW <= 256 && hp.bhr_x <= 256, "dpc_stem_conv_wgrad_s2d: frame too wide (%d)", W);
    hp.Ho = Ho; hp.Wo = Wo; hp.T = T;
    hp.tiles_per_frame = (Ho * hp.PW + 127) / 128;
    const long long total = (long long)NB * T * hp.tiles_per_frame;
    DPC_REQUIRE(total < (1ll << 31), "dpc_stem_conv_wgrad_s2d: too many tiles");
    hp.total_tiles = (int)total;
    hp.xpatch_bytes = ((hp.bhr_x * hp.PW * 32 + 1023) / 1024) * 1024;
    hp.ypatch_bytes = ((hp.bhr_y * hp.PW * 128 + 1023) / 1024) * 1024;
    hp.chain = 64;              // 64 tiles x 8 k-steps x 2 products: 1024 truncating accumulations per TMEM chain
    const size_t smem = 2 * (2 * (size_t)hp.xpatch_bytes + 2 * (size_t)hp.ypatch_bytes) + 64 + 1024;
    DPC_REQUIRE(smem <= 227 * 1024, "dpc_stem_conv_wgrad_s2d: patches do not fit shared memory (W = %d)", W);
    DPC_REQUIRE((hp.ypatch_bytes >> 4) < (1 << 14), "dpc_stem_conv_wgrad_s2d: plane pitch exceeds the descriptor LBO field");
    auto enc = s2d_encode();
    DPC_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled entry point not available");
    S2dWgMaps maps;
    const cuuint32_t es[5] = {1, 1, 1, 1, 1};
    {
        const cuuint64_t gd[5] = {S2D_CH, (cuuint64_t)Wo, (cuuint64_t)Ho, (cuuint64_t)T, (cuuint64_t)NB};
        const cuuint64_t gs[4] = {32, (cuuint64_t)Wo * 32, (cuuint64_t)Ho * Wo * 32, (cuuint64_t)T * Ho * Wo * 32};
        const cuuint32_t bx[5] = {S2D_CH, (cuuint32_t)hp.PW, (cuuint32_t)hp.bhr_x, 1, 1};
        for (int i = 0; i < 2; ++i) {
            CUresult r = enc(i ? &maps.x_lo : &maps.x_hi, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(i ? x2_lo : x2_hi),
                             gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_32B,
                             CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            DPC_REQUIRE(r == CUDA_SUCCESS, "dpc_stem_conv_wgrad_s2d: cuTensorMapEncodeTiled (x) failed (%d)", (int)r);
        }
    }
    {
        const cuuint64_t gd[5] = {64, (cuuint64_t)Wo, (cuuint64_t)Ho, (cuuint64_t)T, (cuuint64_t)NB};
        const cuuint64_t gs[4] = {128, (cuuint64_t)Wo * 128, (cuuint64_t)Ho * Wo * 128, (cuuint64_t)T * Ho * Wo * 128};
        const cuuint32_t bx[5] = {64, (cuuint32_t)hp.PW, (cuuint32_t)hp.bhr_y, 1, 1};
        for (int i = 0; i < 2; ++i) {
            CUresult r = enc(i ? &maps.y_lo : &maps.y_hi, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(i ? dy_lo : dy_hi),
                             gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                             CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            DPC_REQUIRE(r == CUDA_SUCCESS, "dpc_stem_conv_wgrad_s2d: cuTensorMapEncodeTiled (dy) failed (%d)", (int)r);
        }
    }
    DPC_CUDA(cudaMemsetAsync(dw, 0, sizeof(float) * 64 * 147, st));
    DPC_CUDA(cudaFuncSetAttribute(stem_s2d_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int sms = dpc_num_sms();
    const int grid = hp.total_tiles < sms ? hp.total_tiles : sms;
    stem_s2d_wgrad_kernel<<<grid, 192, smem, st>>>(maps, hp, dw);
    DPC_LAUNCH_CHECK();
    return DPC_OK;
}

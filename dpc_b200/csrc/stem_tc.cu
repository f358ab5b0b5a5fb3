// Stem on tcgen05: Conv3d(3,64,(1,7,7),s(1,2,2),p(0,3,3)) forward (+ fused BatchNorm statistics) and
// weight gradient, reading the caller's NCDHW fp32 video directly.
//
// Cin = 3 is too thin for TMA boxes, so the im2col tile [128 output pixels x K] is BUILT in shared
// memory by the CUDA cores.  The fp32 input patch of a tile (3 x 21 x 37) is converted ONCE to bf16
// hi/lo planes; K is laid out as (c, kh, kw padded 7 -> 8), so one 8-element k-group of a pixel is 8
// consecutive patch values: the build is pure 4-byte shared-memory gathers + 16-byte stores into the
// un-swizzled 8x16-byte core-matrix layout (the padded kw = 7 column meets a zero weight).
// The same physical tile is read by the tensor core as
//   * the K-major A operand of the forward GEMM   y[pix, co]  = sum_k  tile[pix, k] * w[co, k]
//   * the MN-major B operand of the wgrad GEMM    dw[co, k]   = sum_pix dy[pix, co] * tile[pix, k]
// (an MN-major view just swaps the descriptor's leading/stride offsets).
// 3xBF16 split, two TMEM accumulators (main + cross terms) like conv_tc.cu.
// Replaces backbone/resnet_2d3d.py:211,260 (self.conv1) forward and its weight gradient.
#include "tc_common.cuh"

namespace {

constexpr int TH = 8, TW = 16;                    // output tile: 8 x 16 = 128 pixels = one UMMA M
constexpr int PH = 2 * TH + 5, PW = 2 * TW + 5;   // input patch 21 x 37 (per channel)
constexpr int PWP = 38;                           // padded patch row (even: 4-byte aligned bf16 pairs)
constexpr int SEG = 21;                           // (c, kh) segments of 8 k-values (kw 0..6 + pad)
constexpr int KG = 24, KP = KG * 8;               // k-groups (21 real + 3 zero), padded depth 192
constexpr int PATCH = 3 * PH * PW;                // 2331 values
constexpr int PATCHP = 3 * PH * PWP;              // 2394 bf16 per plane
constexpr uint32_t A_SBO = KG * 128 + 16;         // bytes between 8-pixel groups (+16: spreads banks)
constexpr uint32_t A_BYTES = 16 * A_SBO;          // 128 pixels
constexpr uint32_t B_SBO = KG * 128;              // forward weights tile: bytes between 8-filter groups
constexpr uint32_t B_BYTES = 8 * B_SBO;           // 64 filters
constexpr uint32_t PLANE_BYTES = ((PATCHP * 2 + 127) / 128) * 128;
constexpr int NPF = (PATCH + 255) / 256;          // patch values prefetched per thread (10)

// shared-memory maps (bytes from a 1024-aligned base)
constexpr uint32_t OFF_A_HI = 0, OFF_A_LO = A_BYTES;
constexpr uint32_t OFF_PL_HI = 2 * A_BYTES, OFF_PL_LO = OFF_PL_HI + PLANE_BYTES;
constexpr uint32_t OFF_COMMON_END = OFF_PL_LO + PLANE_BYTES;
// forward
constexpr uint32_t F_OFF_B_HI = OFF_COMMON_END, F_OFF_B_LO = F_OFF_B_HI + B_BYTES;
constexpr uint32_t F_OFF_STAT = F_OFF_B_LO + B_BYTES;
constexpr uint32_t F_OFF_BAR = F_OFF_STAT + 512;
constexpr uint32_t FWD_SMEM = F_OFF_BAR + 64 + 1024;
// wgrad: dy tiles are TMA boxes (128B swizzle: 1024-byte aligned)
constexpr uint32_t W_OFF_DY_HI = ((OFF_COMMON_END + 1023) / 1024) * 1024;
constexpr uint32_t W_OFF_DY_LO = W_OFF_DY_HI + 16384;
constexpr uint32_t W_OFF_ZERO = W_OFF_DY_LO + 16384;         // second (empty) 64-channel group of A
constexpr uint32_t W_OFF_ACC = W_OFF_ZERO + 16384;           // fp32 [64][KP]
constexpr uint32_t W_OFF_BAR = W_OFF_ACC + 64 * KP * 4;
constexpr uint32_t WGRAD_SMEM = W_OFF_BAR + 64 + 1024;
static_assert(WGRAD_SMEM <= 227 * 1024, "stem wgrad shared memory");

__device__ __forceinline__ uint32_t f2bf_rn(float f) {       // round-to-nearest-even bf16 bits (finite inputs)
    uint32_t u = __float_as_uint(f);
    return (u + 0x7FFFu + ((u >> 16) & 1u)) >> 16;
}

struct TileCoord { int n, t, ho0, wo0; };
__device__ __forceinline__ TileCoord tile_coord(int tile, int T, int tiles_h, int tiles_w) {
    TileCoord c;
    c.wo0 = (tile % tiles_w) * TW; tile /= tiles_w;
    c.ho0 = (tile % tiles_h) * TH; tile /= tiles_h;
    c.t = tile % T;
    c.n = tile / T;
    return c;
}

// issue this thread's share of the NEXT tile's patch loads (kept in registers across the MMA/epilogue)
__device__ __forceinline__ void prefetch_patch(float (&pf)[NPF], const float* __restrict__ x, TileCoord c, int T, int H, int W) {
#pragma unroll
    for (int j = 0; j < NPF; ++j) {
        const int i = threadIdx.x + j * 256;
        float v = 0.f;
        if (i < PATCH) {
            const int col = i % PW, r = (i / PW) % PH, ch = i / (PW * PH);
            const int hi = 2 * c.ho0 - 3 + r, wi = 2 * c.wo0 - 3 + col;
            if (hi >= 0 && hi < H && wi >= 0 && wi < W) v = __ldg(x + ((((size_t)c.n * 3 + ch) * T + c.t) * H + hi) * W + wi);
        }
        pf[j] = v;
    }
}
// registers -> bf16 hi/lo patch planes [3][PH][PWP]
__device__ __forceinline__ void store_patch_planes(const float (&pf)[NPF], uint8_t* smem) {
    unsigned short* ph = reinterpret_cast<unsigned short*>(smem + OFF_PL_HI);
    unsigned short* pl = reinterpret_cast<unsigned short*>(smem + OFF_PL_LO);
#pragma unroll
    for (int j = 0; j < NPF; ++j) {
        const int i = threadIdx.x + j * 256;
        if (i < PATCH) {
            const int col = i % PW, r = i / PW;             // r = ch * PH + row
            const uint32_t h = f2bf_rn(pf[j]);
            ph[r * PWP + col] = (unsigned short)h;
            pl[r * PWP + col] = (unsigned short)f2bf_rn(pf[j] - __uint_as_float(h << 16));
        }
    }
}
// every thread: pixel p = tid % 128, segments [half*11, ...) with half = tid / 128; pure copies
__device__ __forceinline__ void build_im2col_tile(uint8_t* smem) {
    const int p = threadIdx.x & 127, half = threadIdx.x >> 7;
    const int py = p >> 4, px = p & 15;
    uint8_t* rowbase = smem + (uint32_t)(p >> 3) * A_SBO + (uint32_t)(p & 7) * 16;
    const int s_beg = half ? 11 : 0, s_end = half ? SEG : 11;
    for (int seg = s_beg; seg < s_end; ++seg) {             // seg = c * 7 + kh
        const int c = seg / 7, kh = seg - c * 7;
        const uint32_t e = (uint32_t)((c * PH + 2 * py + kh) * PWP + 2 * px) * 2;    // byte offset, 4-aligned
        const uint32_t* sh = reinterpret_cast<const uint32_t*>(smem + OFF_PL_HI + e);
        const uint32_t* sl = reinterpret_cast<const uint32_t*>(smem + OFF_PL_LO + e);
        *reinterpret_cast<uint4*>(rowbase + OFF_A_HI + seg * 128) = make_uint4(sh[0], sh[1], sh[2], sh[3]);
        *reinterpret_cast<uint4*>(rowbase + OFF_A_LO + seg * 128) = make_uint4(sl[0], sl[1], sl[2], sl[3]);
    }
}

// 31-shuffle transposing butterfly: every lane passes 32 per-row values; lane l returns the sum over the
// warp's 32 rows of column l
__device__ __forceinline__ float warp_colsum32(float (&s)[32], int lane) {
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
        const bool up = (lane & off) != 0;
#pragma unroll
        for (int i = 0; i < off; ++i) {
            const float send = up ? s[i] : s[i + off], keep = up ? s[i + off] : s[i];
            s[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
        }
    }
    return s[0];
}

// ================================================================================================
__global__ void __launch_bounds__(256, 1)
stem_tc_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w, float* __restrict__ y,
                   double* __restrict__ stats, int NB, int T, int H, int W, int Ho, int Wo, int total_tiles) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const uint32_t sbase = smem_u32(smem);
    float* stat_smem = reinterpret_cast<float*>(smem + F_OFF_STAT);
    const uint32_t bar = sbase + F_OFF_BAR, tmem_ptr_addr = sbase + F_OFF_BAR + 8;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) tmem_alloc(tmem_ptr_addr, 128);            // 2 accumulators x 64 columns
    // zero the im2col tiles, the patch planes and the weight tiles once: the pad k-groups / pad columns
    // are never written again and must read as finite zeros
    for (uint32_t i = threadIdx.x; i < F_OFF_STAT / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
    if (threadIdx.x < 128) stat_smem[threadIdx.x] = 0.f;
    __syncthreads();
    // weights -> B tiles (K-major core matrices), k = (c*7+kh)*8 + kw, once per CTA
    for (int idx = threadIdx.x; idx < 64 * 147; idx += blockDim.x) {
        const int co = idx / 147, kk = idx % 147;
        const int kw = kk % 7, seg = kk / 7;
        const int k = seg * 8 + kw;
        const float v = w[co * 147 + kk];
        const uint32_t hb = f2bf_rn(v), lb = f2bf_rn(v - __uint_as_float(hb << 16));
        const uint32_t o = (uint32_t)(co >> 3) * B_SBO + (uint32_t)(k >> 3) * 128 + (uint32_t)(co & 7) * 16 + (uint32_t)(k & 7) * 2;
        *reinterpret_cast<unsigned short*>(smem + F_OFF_B_HI + o) = (unsigned short)hb;
        *reinterpret_cast<unsigned short*>(smem + F_OFF_B_LO + o) = (unsigned short)lb;
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_d = *reinterpret_cast<const uint32_t*>(smem + F_OFF_BAR + 8);
    const uint32_t tmem_c = tmem_d + 64;
    // D = f32, A = B = bf16, both K-major, N = 64, M = 128
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((64u >> 3) << 17) | ((128u >> 4) << 24);

    const int tiles_w = (Wo + TW - 1) / TW, tiles_h = (Ho + TH - 1) / TH;
    uint32_t phase = 0;
    float pf[NPF];
    if ((int)blockIdx.x < total_tiles) prefetch_patch(pf, x, tile_coord(blockIdx.x, T, tiles_h, tiles_w), T, H, W);
    int iter = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++iter) {
        const TileCoord tc = tile_coord(tile, T, tiles_h, tiles_w);
        store_patch_planes(pf, smem);
        __syncthreads();
        build_im2col_tile(smem);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
        if (tile + (int)gridDim.x < total_tiles)              // next tile's global loads fly during MMA + epilogue
            prefetch_patch(pf, x, tile_coord(tile + gridDim.x, T, tiles_h, tiles_w), T, H, W);
        if (warp == 0) {
            if (elect_one()) {
                tc_fence_after();
#pragma unroll
                for (int ks = 0; ks < KP / 16; ++ks) {       // UMMA_K = 16 = two k-groups = 256 bytes
                    const uint64_t ahi = make_noswizzle_desc(sbase + OFF_A_HI + ks * 256, 128, A_SBO);
                    const uint64_t alo = make_noswizzle_desc(sbase + OFF_A_LO + ks * 256, 128, A_SBO);
                    const uint64_t bhi = make_noswizzle_desc(sbase + F_OFF_B_HI + ks * 256, 128, B_SBO);
                    const uint64_t blo = make_noswizzle_desc(sbase + F_OFF_B_LO + ks * 256, 128, B_SBO);
                    umma_bf16(tmem_d, ahi, bhi, idesc, ks ? 1u : 0u);
                    umma_bf16(tmem_c, ahi, blo, idesc, ks ? 1u : 0u);
                    umma_bf16(tmem_c, alo, bhi, idesc, 1u);
                }
                umma_commit(bar);
            }
            __syncwarp();
        }
        // epilogue: 8 warps = 4 TMEM lane quarters x 2 column halves
        {
            const int q = warp & 3, ch = warp >> 2;
            const int r = q * 32 + lane;
            const int ho = tc.ho0 + (r >> 4), wo = tc.wo0 + (r & 15);
            const bool valid = ho < Ho && wo < Wo;
            mbar_wait(bar, phase);
            tc_fence_after();
            uint32_t v[32], u[32];
            tmem_ld32(tmem_d + ((uint32_t)(q * 32) << 16) + (uint32_t)(ch * 32), v);
            tmem_ld32(tmem_c + ((uint32_t)(q * 32) << 16) + (uint32_t)(ch * 32), u);
            float s[32], sq[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                const float f = __uint_as_float(v[j]) + __uint_as_float(u[j]);
                v[j] = __float_as_uint(f);
                s[j] = valid ? f : 0.f;
                sq[j] = s[j] * s[j];
            }
            if (valid) {
                float4* dst = reinterpret_cast<float4*>(y + ((((size_t)tc.n * T + tc.t) * Ho + ho) * Wo + wo) * 64 + ch * 32);
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    dst[j] = make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]),
                                         __uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3]));
            }
            if (stats) {
                const float cs = warp_colsum32(s, lane), cq = warp_colsum32(sq, lane);
                atomicAdd(&stat_smem[ch * 32 + lane], cs);
                atomicAdd(&stat_smem[64 + ch * 32 + lane], cq);
            }
        }
        phase ^= 1u;
        tc_fence_before();
        __syncthreads();                 // tile + TMEM free for the next iteration
        tc_fence_after();
        if (stats && (iter & 63) == 63) {                   // flush fp32 partials into fp64 every 64 tiles
            if (threadIdx.x < 128) {
                atomicAdd(stats + threadIdx.x, (double)stat_smem[threadIdx.x]);
                stat_smem[threadIdx.x] = 0.f;
            }
            __syncthreads();
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_d, 128);
    if (stats && threadIdx.x < 128) atomicAdd(stats + threadIdx.x, (double)stat_smem[threadIdx.x]);
}

// ================================================================================================
// dw[co][c][kh][kw] = sum over pixels dy[pix, co] * x_patch[pix, (c,kh,kw)]
constexpr int DRAIN_TILES = 32;      // TMEM chain length between drains: 32 tiles * 8 UMMA steps

__global__ void __launch_bounds__(256, 1)
stem_tc_wgrad_kernel(const __grid_constant__ CUtensorMap map_dy_hi, const __grid_constant__ CUtensorMap map_dy_lo,
                     const float* __restrict__ x, float* __restrict__ dw, int NB, int T, int H, int W, int Ho, int Wo,
                     int total_tiles) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const uint32_t sbase = smem_u32(smem);
    float* acc = reinterpret_cast<float*>(smem + W_OFF_ACC);
    const uint32_t bar_mma = sbase + W_OFF_BAR, bar_dy = sbase + W_OFF_BAR + 8, tmem_ptr_addr = sbase + W_OFF_BAR + 16;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_dy_hi) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_dy_lo) : "memory");
        mbar_init(bar_mma, 1);
        mbar_init(bar_dy, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) tmem_alloc(tmem_ptr_addr, 512);            // 2 accumulators x 192 columns
    for (uint32_t i = threadIdx.x; i < W_OFF_BAR / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_d = *reinterpret_cast<const uint32_t*>(smem + W_OFF_BAR + 16);
    const uint32_t tmem_c = tmem_d + KP;
    // D = f32, A = B = bf16, both MN-major (bits 15, 16), N = 192, M = 128
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(KP >> 3) << 17) |
                           ((128u >> 4) << 24);

    const int tiles_w = (Wo + TW - 1) / TW, tiles_h = (Ho + TH - 1) / TH;
    uint32_t ph_mma = 0, ph_dy = 0;
    float pf[NPF];
    if ((int)blockIdx.x < total_tiles) prefetch_patch(pf, x, tile_coord(blockIdx.x, T, tiles_h, tiles_w), T, H, W);
    int in_chain = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const TileCoord tc = tile_coord(tile, T, tiles_h, tiles_w);
        if (threadIdx.x == 0) {                               // dy tile: [8 x 16 pixels] x 64 channels, hi and lo
            mbar_expect_tx(bar_dy, 2u * 16384u);
            tma_load_5d(&map_dy_hi, sbase + W_OFF_DY_HI, bar_dy, 0, tc.wo0, tc.ho0, tc.t, tc.n);
            tma_load_5d(&map_dy_lo, sbase + W_OFF_DY_LO, bar_dy, 0, tc.wo0, tc.ho0, tc.t, tc.n);
        }
        store_patch_planes(pf, smem);
        __syncthreads();
        build_im2col_tile(smem);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
        if (tile + (int)gridDim.x < total_tiles)
            prefetch_patch(pf, x, tile_coord(tile + gridDim.x, T, tiles_h, tiles_w), T, H, W);
        if (warp == 0) {
            if (elect_one()) {
                mbar_wait(bar_dy, ph_dy);
                tc_fence_after();
#pragma unroll
                for (int ks = 0; ks < 8; ++ks) {             // UMMA_K = 16 pixels
                    // A = dy^T: MN-major 128B-swizzled TMA box, 16 rows = 2 atoms = 2048 B per step; the second
                    // 64-channel group (M rows 64..127) is a zero region
                    const uint64_t ahi = make_mnmajor_sw128_desc(sbase + W_OFF_DY_HI + ks * 2048, W_OFF_ZERO - W_OFF_DY_HI);
                    const uint64_t alo = make_mnmajor_sw128_desc(sbase + W_OFF_DY_LO + ks * 2048, W_OFF_ZERO - W_OFF_DY_LO);
                    // B = im2col tile read MN-major: 8-pixel groups (K) are A_SBO apart, 8-value k-groups (N) 128 B apart
                    const uint64_t bhi = make_noswizzle_desc(sbase + OFF_A_HI + ks * 2 * A_SBO, A_SBO, 128);
                    const uint64_t blo = make_noswizzle_desc(sbase + OFF_A_LO + ks * 2 * A_SBO, A_SBO, 128);
                    const uint32_t first = (in_chain | ks) ? 1u : 0u;
                    umma_bf16(tmem_d, ahi, bhi, idesc, first);
                    umma_bf16(tmem_c, ahi, blo, idesc, first);
                    umma_bf16(tmem_c, alo, bhi, idesc, 1u);
                }
                umma_commit(bar_mma);
            }
            __syncwarp();
        }
        ph_dy ^= 1u;
        ++in_chain;
        const bool last = tile + (int)gridDim.x >= total_tiles;
        // all threads wait for the MMAs before the tiles are overwritten
        mbar_wait(bar_mma, ph_mma);
        ph_mma ^= 1u;
        tc_fence_after();
        if (in_chain == DRAIN_TILES || last) {
            // drain the TMEM chain into the fp32 shared accumulators (round-to-nearest adds)
            if (warp < 2) {                                   // TMEM lanes 0..63 = output channels
                const int co = warp * 32 + lane;
                for (int c0 = 0; c0 < KP; c0 += 32) {
                    uint32_t v[32], u[32];
                    tmem_ld32(tmem_d + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, v);
                    tmem_ld32(tmem_c + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, u);
#pragma unroll
                    for (int j = 0; j < 32; ++j) acc[co * KP + c0 + j] += __uint_as_float(v[j]) + __uint_as_float(u[j]);
                }
            }
            in_chain = 0;
        }
        tc_fence_before();
        __syncthreads();
        tc_fence_after();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_d, 512);
    // shared accumulators -> global (skip the padded kw = 7 column and the zero k-groups)
    for (int idx = threadIdx.x; idx < 64 * 147; idx += blockDim.x) {
        const int co = idx / 147, kk = idx % 147;
        const int k = (kk / 7) * 8 + kk % 7;
        atomicAdd(dw + idx, acc[co * KP + k]);
    }
}

}  // namespace

// y [NB,T,Ho,Wo,64] = conv1(x [NB,3,T,H,W]); bn_ws (nullable): 128 doubles = per-channel sum | sum of squares
extern "C" int dpc_stem_conv_fwd_tc(const float* x, const float* w, float* y, double* bn_ws, int NB, int T, int H, int W,
                                    void* stream) {
    DPC_REQUIRE(x && w && y && NB > 0 && T > 0 && H > 0 && W > 0, "dpc_stem_conv_fwd_tc: bad args");
    const int Ho = (H + 6 - 7) / 2 + 1, Wo = (W + 6 - 7) / 2 + 1;
    const long long tiles = (long long)NB * T * ((Ho + TH - 1) / TH) * ((Wo + TW - 1) / TW);
    DPC_REQUIRE(tiles < (1ll << 31), "dpc_stem_conv_fwd_tc: too many tiles");
    cudaStream_t st = as_stream(stream);
    if (bn_ws) DPC_CUDA(cudaMemsetAsync(bn_ws, 0, sizeof(double) * 128, st));
    DPC_CUDA(cudaFuncSetAttribute(stem_tc_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FWD_SMEM));
    int grid = dpc_num_sms();
    if (grid > tiles) grid = (int)tiles;
    stem_tc_fwd_kernel<<<grid, 256, FWD_SMEM, st>>>(x, w, y, bn_ws, NB, T, H, W, Ho, Wo, (int)tiles);
    DPC_LAUNCH_CHECK();
    return DPC_OK;
}

// dw [64,3,1,7,7] = wgrad(x [NB,3,T,H,W], dy planes [NB,T,Ho,Wo,64] split-bf16)
extern "C" int dpc_stem_conv_wgrad_tc(const float* x, const void* dy_hi, const void* dy_lo, float* dw, int NB, int T,
                                      int H, int W, void* stream) {
    DPC_REQUIRE(x && dy_hi && dy_lo && dw && NB > 0 && T > 0 && H > 0 && W > 0, "dpc_stem_conv_wgrad_tc: bad args");
    const int Ho = (H + 6 - 7) / 2 + 1, Wo = (W + 6 - 7) / 2 + 1;
    const long long tiles = (long long)NB * T * ((Ho + TH - 1) / TH) * ((Wo + TW - 1) / TW);
    DPC_REQUIRE(tiles < (1ll << 31), "dpc_stem_conv_wgrad_tc: too many tiles");
    cudaStream_t st = as_stream(stream);
    static PFN_cuTensorMapEncodeTiled_v12000 enc = nullptr;
    if (!enc) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            enc = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
    }
    DPC_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled entry point not available");
    CUtensorMap mh, ml;
    const cuuint64_t gd[5] = {64, (cuuint64_t)Wo, (cuuint64_t)Ho, (cuuint64_t)T, (cuuint64_t)NB};
    const cuuint64_t gs[4] = {128, (cuuint64_t)Wo * 128, (cuuint64_t)Ho * Wo * 128, (cuuint64_t)T * Ho * Wo * 128};
    const cuuint32_t bx[5] = {64, TW, TH, 1, 1}, es[5] = {1, 1, 1, 1, 1};
    for (int i = 0; i < 2; ++i) {
        CUresult r = enc(i ? &ml : &mh, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(i ? dy_lo : dy_hi), gd, gs, bx,
                         es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        DPC_REQUIRE(r == CUDA_SUCCESS, "dpc_stem_conv_wgrad_tc: cuTensorMapEncodeTiled failed (%d)", (int)r);
    }
    DPC_CUDA(cudaMemsetAsync(dw, 0, sizeof(float) * 64 * 147, st));
    DPC_CUDA(cudaFuncSetAttribute(stem_tc_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)WGRAD_SMEM));
    int grid = dpc_num_sms();
    if (grid > tiles) grid = (int)tiles;
    stem_tc_wgrad_kernel<<<grid, 256, WGRAD_SMEM, st>>>(mh, ml, x, dw, NB, T, H, W, Ho, Wo, (int)tiles);
    DPC_LAUNCH_CHECK();
    return DPC_OK;
}

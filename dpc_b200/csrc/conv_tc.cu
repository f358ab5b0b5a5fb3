// tcgen05 (5th-gen tensor core) implicit-GEMM kernels for the 1x3x3 / 3x3x3 / 1x1x1 convolutions
// (forward, dgrad, wgrad; any stride) and the dense score matmul.  sm_100a only.
//
//   D (fp32, TMEM) += sum over k-blocks  A_hi*B_hi + A_hi*B_lo + A_lo*B_hi          (3xBF16 split)
//
// * operands live in HBM as two bf16 planes (hi = bf16(x), lo = bf16(x - hi)): same bytes as fp32,
//   ~16 mantissa bits -- what the 1e-3 parity bar needs (1-pass BF16/TF32 fail it).  fp16 pairs
//   (22 bits) would suit the O(1) forward values, but tcgen05 kind::f16 rejects mixed A / B formats (measured: an
//   fp16 x bf16 descriptor raises "illegal instruction"), wgrad multiplies activations by gradients, and gradients
//   need bf16's exponent range -- so activations stay bf16 pairs; only GEMMs whose operands are BOTH forward values
//   (the score matmul) take fp16 pairs (dpc_gemm_nt_split_tc);
// * an operand tile (positions x 64 channels of ONE filter tap) is ONE TMA box of the channels-last
//   activation tensor [NB,T,H,W,C] at the tap-shifted coordinate; halo / zero padding is TMA
//   out-of-bounds fill: no im2col buffer, no index arithmetic on the SM.  Strided convolutions read
//   through per-parity views of the tensor (base pointer offset + doubled strides in the tensor map);
// * shared-memory tiles are 128B-swizzled; forward/dgrad use K-major UMMA descriptors (K = channels),
//   wgrad uses MN-major descriptors over the SAME boxes (K = positions);
// * warp 0 = TMA producer, warp 1 = MMA issuer (one elected lane) + TMEM owner, warps 2-5 (2-9 in the halo kernel)
//   = epilogue (tcgen05.ld -> registers -> lane-pair exchange -> full-sector stores), with an mbarrier full/empty
//   ring between them;
// * three schedules: conv_tc_kernel (one tile per CTA, 2 CTAs/SM, any shape), conv_tc_persist_kernel (Co <= 128:
//   persistent, double-buffered TMEM) and the halo-patch kernels conv_tc_halo_kernel / wgrad_halo_kernel for the
//   stride-1 1x3x3 64 -> 64 sites (one TMA box per tile, taps = row-shifted descriptors); see DESIGN.md section 5.
//
// Replaces nn.Conv3d fwd/bwd at backbone/resnet_2d3d.py:13-31,241-244 and torch.matmul at
// dpc/model_3d.py:83.
#include "tc_common.cuh"
#include <cuda_fp16.h>
#include <stdlib.h>

namespace {

// One filter-tap table entry per dimension: which original tap, which coordinate offset of the
// gathered tensor relative to the tile origin, which parity view (strided convs).
struct TapDim {
    int8_t count;
    int8_t k[3];       // original tap index along this dimension
    int8_t off[3];     // coordinate offset
    int8_t par[3];     // parity class (0..stride-1)
};

struct TcMaps {
    CUtensorMap a_hi[8], a_lo[8];     // gathered tensor, one per parity view
    CUtensorMap b_hi, b_lo;           // second operand
};

// Optional fused reduction of the epilogue (sums over output rows per output channel, written to `stats` as
// [sum_a (Co) | sum_b (Co)] doubles):
//   y == nullptr: BatchNorm batch statistics of the conv output v:      a = v,  b = v^2           (forward)
//   y != nullptr: BatchNorm-BACKWARD sums of the BN that produced the conv's INPUT gradient target, i.e. with
//                 g = v * [mask_hi > 0] (ReLU mask of that BN's output; nullptr = no ReLU) and
//                 xhat = (y - mean) * rstd (that BN's input):            a = g,  b = g * xhat      (dgrad)
// so bn_bwd's separate reduce pass over dgrad's output disappears (bn.cu: dpc_bn_bwd_apply consumes the sums).
struct BnRed {
    const __nv_bfloat16* mask_hi;
    const float* y;
    const float* mean;
    const float* rstd;
};

struct TcParams {
    BnRed red;
    TapDim tT, tH, tW;
    int kH, kW;                // full filter extents (to linearise the original tap index)
    int sH, sW;                // strides (to linearise the parity-view index)
    int cchunks, Ksrc;         // 64-channel chunks / channels of the gathered tensor
    int bw, bh, bt, bn, box_rows;
    int tiles_w, tiles_h, tiles_t, tiles_n;
    int NB, To, Ho, Wo, Co;    // extents of the tile grid and the channel count of the output
    int BN, stages, nviews;
    long long out_sn, out_st, out_sh, out_sw, out_base;   // output row = n*sn + t*st + h*sh + w*sw + base
    // wgrad only
    int taps_full, splits, ktiles_per_split;
    int tap_group;             // taps per CTA (they share the dY tile; N = tap_group * BN)
    int ab_f16;                // K-major kernels: 1 = both operands are fp16 pairs (score matmul), 0 = bf16 pairs
};

constexpr int A_TILE_BYTES = 128 * 128;            // 128 rows x 64 bf16

struct SmemPlan {
    uint32_t base, stage_bytes, b_tile_bytes, bar_base;
    int stages;
    __device__ uint32_t full(int s) const { return bar_base + 8u * s; }
    __device__ uint32_t empty(int s) const { return bar_base + 8u * (stages + s); }
    __device__ uint32_t tmem_full() const { return bar_base + 8u * (2 * stages); }
    __device__ uint32_t tmem_ptr() const { return bar_base + 8u * (2 * stages + 1); }
};

__device__ __forceinline__ SmemPlan plan_smem(const uint8_t* smem_raw, int BN, int stages) {
    SmemPlan s;
    s.base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    s.b_tile_bytes = (uint32_t)BN * 128u;
    s.stage_bytes = 2u * A_TILE_BYTES + 2u * s.b_tile_bytes;
    s.stages = stages;
    s.bar_base = s.base + (uint32_t)stages * s.stage_bytes;
    return s;
}

__device__ __forceinline__ uint32_t setup_common(const SmemPlan& sp, const uint8_t* smem_raw, const TcMaps& maps,
                                                 int nviews, uint32_t tmem_cols) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0 && lane == 0) {
        for (int v = 0; v < nviews; ++v) {
            asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.a_hi[v]) : "memory");
            asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.a_lo[v]) : "memory");
        }
        asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.b_hi) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.b_lo) : "memory");
        for (int s = 0; s < sp.stages; ++s) { mbar_init(sp.full(s), 1); mbar_init(sp.empty(s), 1); }
        mbar_init(sp.tmem_full(), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) tmem_alloc(sp.tmem_ptr(), tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    return *reinterpret_cast<const uint32_t*>(smem_raw + (sp.tmem_ptr() - smem_u32(smem_raw)));
}

// One warp's share of a finished tile: TMEM lanes [32q, 32q+32) = tile rows; adds the two accumulators,
// stores the fp32 rows, and (optionally) accumulates the BatchNorm partial sums of the tile.
__device__ __forceinline__ void tile_epilogue_rows(const TcParams& p, float* __restrict__ y, int accumulate,
                                                   float* stat_smem, uint32_t tmem_d, uint32_t tmem_c, int q, int lane,
                                                   bool valid, long long row, int ncol0);

__device__ __forceinline__ void tile_epilogue(const TcParams& p, float* __restrict__ y, int accumulate,
                                              float* stat_smem, uint32_t tmem_d, uint32_t tmem_c, int q, int lane,
                                              int n0, int t0, int h0, int w0, int ncol0) {
    const int r = q * 32 + lane;
    const int dw = r % p.bw, dh = (r / p.bw) % p.bh, dt = (r / (p.bw * p.bh)) % p.bt, dn = r / (p.bw * p.bh * p.bt);
    const int n = n0 + dn, t = t0 + dt, h = h0 + dh, w = w0 + dw;
    const bool valid = r < p.box_rows && n < p.NB && t < p.To && h < p.Ho && w < p.Wo;
    const long long row = (long long)n * p.out_sn + (long long)t * p.out_st + (long long)h * p.out_sh +
                          (long long)w * p.out_sw + p.out_base;
    tile_epilogue_rows(p, y, accumulate, stat_smem, tmem_d, tmem_c, q, lane, valid, row, ncol0);
}

// `valid` / `row`: whether this lane's accumulator row is a real output position, and which one
__device__ __forceinline__ void tile_epilogue_rows(const TcParams& p, float* __restrict__ y, int accumulate,
                                                   float* stat_smem, uint32_t tmem_d, uint32_t tmem_c, int q, int lane,
                                                   bool valid, long long row, int ncol0) {
    float* yrow = y + (valid ? row : 0) * p.Co + ncol0;
    const bool vec = (p.Co & 3) == 0;
    for (int c0 = 0; c0 < p.BN; c0 += 32) {
        uint32_t v[32], u[32];
        tmem_ld32_nowait(tmem_d + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
        tmem_ld32_nowait(tmem_c + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, u);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) + __uint_as_float(u[j]));
        const bool fast = vec && ncol0 + c0 + 32 <= p.Co && !p.red.y;      // CTA-uniform
        if (fast) {
            // full-sector stores through lane pairs (tc_common.cuh); v keeps this lane's own pre-accumulate row
            float4 o[8];
#pragma unroll
            for (int j = 0; j < 8; ++j)
                o[j] = make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]), __uint_as_float(v[4 * j + 2]),
                                   __uint_as_float(v[4 * j + 3]));
            pair_store_rows<4>(o, y, row, valid, lane, p.Co, ncol0 + c0, accumulate != 0);
        } else if (valid) {
            if (vec && ncol0 + c0 + 32 <= p.Co) {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    float4 o = make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]),
                                           __uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3]));
                    float4* dst = reinterpret_cast<float4*>(yrow + c0) + j;
                    if (accumulate) {
                        float4 c = *dst; o.x += c.x; o.y += c.y; o.z += c.z; o.w += c.w;
                        v[4 * j] = __float_as_uint(o.x); v[4 * j + 1] = __float_as_uint(o.y);
                        v[4 * j + 2] = __float_as_uint(o.z); v[4 * j + 3] = __float_as_uint(o.w);
                    }
                    *dst = o;
                }
            } else {
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    if (ncol0 + c0 + j < p.Co) {
                        float o = __uint_as_float(v[j]);
                        if (accumulate) { o += yrow[c0 + j]; v[j] = __float_as_uint(o); }
                        yrow[c0 + j] = o;
                    }
                }
            }
        }
        if (stat_smem) {
            // BatchNorm statistics of this tile, fused into the producer: per-column sum / sum of squares
            // over the warp's 32 rows by a 31-shuffle transposing butterfly (lane l ends up with column
            // c0+l), then shared-memory partials per CTA.
            float sv[32], sq[32];
            if (p.red.y) {
                // BatchNorm-backward sums (see BnRed): g = v * [mask > 0], b = g * (y - mean) * rstd; mean / rstd of
                // this CTA's columns sit in shared memory behind the partial sums
                const bool full = valid && ncol0 + c0 + 32 <= p.Co && (p.Co & 7) == 0;
                const size_t off = (size_t)(valid ? row : 0) * p.Co + ncol0 + c0;
#pragma unroll
                for (int j8 = 0; j8 < 4; ++j8) {
                    uint4 mk = make_uint4(0x3f803f80u, 0x3f803f80u, 0x3f803f80u, 0x3f803f80u);      // +1: keep
                    float4 ya = make_float4(0.f, 0.f, 0.f, 0.f), yb = ya;
                    if (full) {
                        if (p.red.mask_hi) mk = *reinterpret_cast<const uint4*>(p.red.mask_hi + off + 8 * j8);
                        ya = *reinterpret_cast<const float4*>(p.red.y + off + 8 * j8);
                        yb = *reinterpret_cast<const float4*>(p.red.y + off + 8 * j8 + 4);
                    }
                    const uint32_t mw[4] = {mk.x, mk.y, mk.z, mk.w};
                    const float yv[8] = {ya.x, ya.y, ya.z, ya.w, yb.x, yb.y, yb.z, yb.w};
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        const int j = 8 * j8 + e;
                        // bf16 > 0  <=>  sign bit clear and not zero
                        const uint32_t hb = (e & 1) ? (mw[e >> 1] >> 16) : (mw[e >> 1] & 0xffffu);
                        const bool keep = full && (hb & 0x8000u) == 0u && (hb & 0x7fffu) != 0u;
                        const float gv = keep ? __uint_as_float(v[j]) : 0.f;
                        sv[j] = gv;
                        sq[j] = gv * (yv[e] - stat_smem[512 + c0 + j]) * stat_smem[768 + c0 + j];
                    }
                }
            } else {
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const float f = valid ? __uint_as_float(v[j]) : 0.f;
                    sv[j] = f; sq[j] = f * f;
                }
            }
#pragma unroll
            for (int off = 16; off >= 1; off >>= 1) {
                const bool up = (lane & off) != 0;
#pragma unroll
                for (int i = 0; i < off; ++i) {
                    const float s_send = up ? sv[i] : sv[i + off], s_keep = up ? sv[i + off] : sv[i];
                    const float q_send = up ? sq[i] : sq[i + off], q_keep = up ? sq[i + off] : sq[i];
                    sv[i] = s_keep + __shfl_xor_sync(0xffffffffu, s_send, off);
                    sq[i] = q_keep + __shfl_xor_sync(0xffffffffu, q_send, off);
                }
            }
            atomicAdd(&stat_smem[c0 + lane], sv[0]);
            atomicAdd(&stat_smem[256 + c0 + lane], sq[0]);
        }
    }
}

// =============================================================================================
// forward / dgrad / plain GEMM:  out[position, co] = sum_{tap, c} G[position (+) tap, c] * Wp[co][tap][c]
// =============================================================================================
__global__ void __launch_bounds__(192, 2)
conv_tc_kernel(const __grid_constant__ TcMaps maps, const TcParams p, float* __restrict__ y, int accumulate,
               double* __restrict__ stats) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const SmemPlan sp = plan_smem(smem_raw, p.BN, p.stages);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // per-CTA BatchNorm partials (sum | sum of squares), after the barrier block
    float* stat_smem = reinterpret_cast<float*>(smem_raw + (sp.bar_base + 8u * (2 * p.stages + 2) - smem_u32(smem_raw)));
    if (stats) {
        for (int i = threadIdx.x; i < 512; i += blockDim.x) stat_smem[i] = 0.f;
        if (p.red.y)
            for (int c = threadIdx.x; c < p.BN; c += blockDim.x) {
                const int col = blockIdx.y * p.BN + c;
                stat_smem[512 + c] = col < p.Co ? p.red.mean[col] : 0.f;
                stat_smem[768 + c] = col < p.Co ? p.red.rstd[col] : 0.f;
            }
    }
    // Two accumulators: columns [0,BN) take hi*hi, [BN,2BN) the two cross terms.  The TMEM accumulate
    // truncates (measured: mean relative error -2e-8 per accumulation step), so keeping the small terms
    // out of the main chain cuts that bias 3x; the epilogue adds the two in fp32 (round-to-nearest).
    const uint32_t tmem_cols = 2u * (p.BN < 16 ? 16 : p.BN);
    const uint32_t tmem_d = setup_common(sp, smem_raw, maps, p.nviews, tmem_cols);
    const uint32_t tmem_c = tmem_d + (uint32_t)p.BN;

    int tile = blockIdx.x;
    const int tw = tile % p.tiles_w; tile /= p.tiles_w;
    const int th = tile % p.tiles_h; tile /= p.tiles_h;
    const int tt = tile % p.tiles_t; tile /= p.tiles_t;
    const int tn = tile;
    const int w0 = tw * p.bw, h0 = th * p.bh, t0 = tt * p.bt, n0 = tn * p.bn;
    const int ncol0 = blockIdx.y * p.BN;
    const int ntaps = p.tT.count * p.tH.count * p.tW.count;
    const int num_kb = ntaps * p.cchunks;

    if (warp == 0) {
        if (elect_one()) {
            const uint32_t tx = 2u * (uint32_t)(p.box_rows * 128) + 2u * sp.b_tile_bytes;
            int s = 0; uint32_t ph = 0;
            for (int kb = 0; kb < num_kb; ++kb) {
                const int tap = kb / p.cchunks, cc = kb - tap * p.cchunks;
                const int iw = tap % p.tW.count, ih = (tap / p.tW.count) % p.tH.count, it = tap / (p.tW.count * p.tH.count);
                const int tap_full = (p.tT.k[it] * p.kH + p.tH.k[ih]) * p.kW + p.tW.k[iw];
                const int view = (p.tT.par[it] * p.sH + p.tH.par[ih]) * p.sW + p.tW.par[iw];
                mbar_wait(sp.empty(s), ph ^ 1u);
                mbar_expect_tx(sp.full(s), tx);
                const uint32_t sa = sp.base + s * sp.stage_bytes;
                const int cw = w0 + p.tW.off[iw], chh = h0 + p.tH.off[ih], ct = t0 + p.tT.off[it];
                tma_load_5d(&maps.a_hi[view], sa, sp.full(s), cc * 64, cw, chh, ct, n0);
                tma_load_5d(&maps.a_lo[view], sa + A_TILE_BYTES, sp.full(s), cc * 64, cw, chh, ct, n0);
                const int kcol = tap_full * p.Ksrc + cc * 64;
                tma_load_2d(&maps.b_hi, sa + 2 * A_TILE_BYTES, sp.full(s), kcol, ncol0);
                tma_load_2d(&maps.b_lo, sa + 2 * A_TILE_BYTES + sp.b_tile_bytes, sp.full(s), kcol, ncol0);
                if (++s == p.stages) { s = 0; ph ^= 1u; }
            }
        }
    } else if (warp == 1) {
        if (elect_one()) {
            // instruction descriptor: D=f32, A=B=bf16, both K-major, N = BN, M = 128
            const uint32_t fmt = (1u << 4) | (p.ab_f16 ? 0u : ((1u << 7) | (1u << 10)));     // D = f32; A = B: bf16 (1) or f16 (0)
            const uint32_t idesc = fmt | ((uint32_t)(p.BN >> 3) << 17) | ((128u >> 4) << 24);
            // the B_hi and B_lo tiles are adjacent in a stage and so are the two accumulators: for BN <= 128 they
            // are one N = 2*BN operand / destination
            const bool wide = p.BN <= 128 && p.BN >= 16;
            const uint32_t idesc2 = fmt | ((uint32_t)((2 * p.BN) >> 3) << 17) | ((128u >> 4) << 24);
            int s = 0; uint32_t ph = 0;
            for (int kb = 0; kb < num_kb; ++kb) {
                mbar_wait(sp.full(s), ph);
                tc_fence_after();
                const uint32_t sa = sp.base + s * sp.stage_bytes;
                const uint64_t ahi = make_kmajor_sw128_desc(sa), alo = make_kmajor_sw128_desc(sa + A_TILE_BYTES);
                const uint64_t bhi = make_kmajor_sw128_desc(sa + 2 * A_TILE_BYTES);
                const uint64_t blo = make_kmajor_sw128_desc(sa + 2 * A_TILE_BYTES + sp.b_tile_bytes);
#pragma unroll
                for (int k = 0; k < 4; ++k) {               // 4 x UMMA_K(16) = 64 channels; +32 B per step
                    const uint64_t ko = (uint64_t)(k * 2);
                    if (wide) {                              // A_hi x [B_hi ; B_lo]: one instruction, both accumulators
                        umma_bf16(tmem_d, ahi + ko, bhi + ko, idesc2, (kb | k) ? 1u : 0u);
                    } else {
                        umma_bf16(tmem_d, ahi + ko, bhi + ko, idesc, (kb | k) ? 1u : 0u);
                        umma_bf16(tmem_c, ahi + ko, blo + ko, idesc, (kb | k) ? 1u : 0u);
                    }
                    umma_bf16(tmem_c, alo + ko, bhi + ko, idesc, 1u);
                }
                umma_commit(sp.empty(s));                   // frees the smem stage when these MMAs retire
                if (++s == p.stages) { s = 0; ph ^= 1u; }
            }
            umma_commit(sp.tmem_full());
        }
    } else {
        // epilogue: warps 2..5; TMEM lane quarter = warp % 4
        mbar_wait(sp.tmem_full(), 0);
        tc_fence_after();
        tile_epilogue(p, y, accumulate, stats ? stat_smem : nullptr, tmem_d, tmem_c, warp & 3, lane, n0, t0, h0, w0, ncol0);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_d, tmem_cols);
    if (stats) {
        for (int c = threadIdx.x; c < p.BN; c += blockDim.x) {
            if (ncol0 + c < p.Co) {
                atomicAdd(stats + ncol0 + c, (double)stat_smem[c]);
                atomicAdd(stats + p.Co + ncol0 + c, (double)stat_smem[256 + c]);
            }
        }
    }
}

// =============================================================================================
// Persistent variant for narrow outputs (BN <= 128, one N tile): each CTA walks many tiles, the TMEM
// accumulators are double-buffered so the epilogue of tile i overlaps the MMAs of tile i+1, and -- when the
// whole packed filter bank fits (layer1: 9 taps x 64 x 64 x hi/lo = 144 KB) -- the weights stay RESIDENT
// in shared memory instead of being re-fetched from L2 for every tile (layer1 is L2-bandwidth bound).
// =============================================================================================
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

__global__ void __launch_bounds__(192, 1)
conv_tc_persist_kernel(const __grid_constant__ TcMaps maps, const TcParams p, float* __restrict__ y, int accumulate,
                       double* __restrict__ stats, int resident_b, int total_tiles) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ntaps = p.tT.count * p.tH.count * p.tW.count;
    const int num_kb = ntaps * p.cchunks;
    const uint32_t b_tile = (uint32_t)p.BN * 128u;
    // smem: [resident weights: num_kb x (hi | lo)] [stages x (A_hi | A_lo [| B_hi | B_lo])] [barriers] [stats]
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t w_bytes = resident_b ? (uint32_t)num_kb * 2u * b_tile : 0u;
    const uint32_t stage_bytes = 2u * A_TILE_BYTES + (resident_b ? 0u : 2u * b_tile);
    const uint32_t ring = base + w_bytes;
    const uint32_t bar_base = ring + (uint32_t)p.stages * stage_bytes;
    auto full = [&](int s) { return bar_base + 8u * s; };
    auto empty = [&](int s) { return bar_base + 8u * (p.stages + s); };
    const uint32_t w_full = bar_base + 8u * (2 * p.stages);
    auto tm_full = [&](int b) { return bar_base + 8u * (2 * p.stages + 1 + b); };
    auto tm_empty = [&](int b) { return bar_base + 8u * (2 * p.stages + 3 + b); };
    const uint32_t tmem_ptr_addr = bar_base + 8u * (2 * p.stages + 5);
    float* stat_smem = reinterpret_cast<float*>(smem_raw + (bar_base + 8u * (2 * p.stages + 6) - smem_u32(smem_raw)));
    if (stats) {
        for (int i = threadIdx.x; i < 512; i += blockDim.x) stat_smem[i] = 0.f;
        if (p.red.y)
            for (int c = threadIdx.x; c < p.BN; c += blockDim.x) {
                stat_smem[512 + c] = c < p.Co ? p.red.mean[c] : 0.f;
                stat_smem[768 + c] = c < p.Co ? p.red.rstd[c] : 0.f;
            }
    }
    const uint32_t tmem_cols = 4u * (uint32_t)p.BN;            // 2 buffers x (main + cross-term) accumulators
    if (warp == 0 && lane == 0) {
        for (int v = 0; v < p.nviews; ++v) {
            asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.a_hi[v]) : "memory");
            asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.a_lo[v]) : "memory");
        }
        asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.b_hi) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.b_lo) : "memory");
        for (int s = 0; s < p.stages; ++s) { mbar_init(full(s), 1); mbar_init(empty(s), 1); }
        mbar_init(w_full, 1);
        for (int b = 0; b < 2; ++b) { mbar_init(tm_full(b), 1); mbar_init(tm_empty(b), 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) tmem_alloc(tmem_ptr_addr, tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<const uint32_t*>(smem_raw + (tmem_ptr_addr - smem_u32(smem_raw)));
    const int my_tiles = ((int)blockIdx.x < total_tiles) ? (total_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;

    auto tile_origin = [&](int i, int& n0, int& t0, int& h0, int& w0) {
        int tile = (int)blockIdx.x + i * (int)gridDim.x;
        const int tw = tile % p.tiles_w; tile /= p.tiles_w;
        const int th = tile % p.tiles_h; tile /= p.tiles_h;
        const int tt = tile % p.tiles_t; tile /= p.tiles_t;
        w0 = tw * p.bw; h0 = th * p.bh; t0 = tt * p.bt; n0 = tile * p.bn;
    };

    if (warp == 0) {
        if (elect_one()) {
            if (resident_b && my_tiles > 0) {
                mbar_expect_tx(w_full, w_bytes);
                for (int kb = 0; kb < num_kb; ++kb) {
                    const int tap = kb / p.cchunks, cc = kb - tap * p.cchunks;
                    const int iw = tap % p.tW.count, ih = (tap / p.tW.count) % p.tH.count, it = tap / (p.tW.count * p.tH.count);
                    const int tap_full = (p.tT.k[it] * p.kH + p.tH.k[ih]) * p.kW + p.tW.k[iw];
                    const int kcol = tap_full * p.Ksrc + cc * 64;
                    tma_load_2d(&maps.b_hi, base + (uint32_t)kb * 2u * b_tile, w_full, kcol, 0);
                    tma_load_2d(&maps.b_lo, base + (uint32_t)kb * 2u * b_tile + b_tile, w_full, kcol, 0);
                }
            }
            const uint32_t tx = 2u * (uint32_t)(p.box_rows * 128) + (resident_b ? 0u : 2u * b_tile);
            int s = 0; uint32_t ph = 0;
            for (int i = 0; i < my_tiles; ++i) {
                int n0, t0, h0, w0;
                tile_origin(i, n0, t0, h0, w0);
                for (int kb = 0; kb < num_kb; ++kb) {
                    const int tap = kb / p.cchunks, cc = kb - tap * p.cchunks;
                    const int iw = tap % p.tW.count, ih = (tap / p.tW.count) % p.tH.count, it = tap / (p.tW.count * p.tH.count);
                    const int view = (p.tT.par[it] * p.sH + p.tH.par[ih]) * p.sW + p.tW.par[iw];
                    mbar_wait(empty(s), ph ^ 1u);
                    mbar_expect_tx(full(s), tx);
                    const uint32_t sa = ring + s * stage_bytes;
                    const int cw = w0 + p.tW.off[iw], chh = h0 + p.tH.off[ih], ct = t0 + p.tT.off[it];
                    tma_load_5d(&maps.a_hi[view], sa, full(s), cc * 64, cw, chh, ct, n0);
                    tma_load_5d(&maps.a_lo[view], sa + A_TILE_BYTES, full(s), cc * 64, cw, chh, ct, n0);
                    if (!resident_b) {
                        const int tap_full = (p.tT.k[it] * p.kH + p.tH.k[ih]) * p.kW + p.tW.k[iw];
                        const int kcol = tap_full * p.Ksrc + cc * 64;
                        tma_load_2d(&maps.b_hi, sa + 2 * A_TILE_BYTES, full(s), kcol, 0);
                        tma_load_2d(&maps.b_lo, sa + 2 * A_TILE_BYTES + b_tile, full(s), kcol, 0);
                    }
                    if (++s == p.stages) { s = 0; ph ^= 1u; }
                }
            }
        }
    } else if (warp == 1) {
        if (elect_one()) {
            const uint32_t fmt = (1u << 4) | (p.ab_f16 ? 0u : ((1u << 7) | (1u << 10)));
            const uint32_t idesc = fmt | ((uint32_t)(p.BN >> 3) << 17) | ((128u >> 4) << 24);
            const uint32_t idesc2 = fmt | ((uint32_t)((2 * p.BN) >> 3) << 17) | ((128u >> 4) << 24);
            if (resident_b && my_tiles > 0) mbar_wait(w_full, 0);
            int s = 0; uint32_t ph = 0;
            for (int i = 0; i < my_tiles; ++i) {
                const int buf = i & 1;
                const uint32_t td = tmem_base + (uint32_t)(buf * 2 * p.BN), tcx = td + (uint32_t)p.BN;
                mbar_wait(tm_empty(buf), (((uint32_t)i >> 1) & 1u) ^ 1u);        // epilogue drained this buffer
                tc_fence_after();
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(full(s), ph);
                    tc_fence_after();
                    const uint32_t sa = ring + s * stage_bytes;
                    const uint32_t sb = resident_b ? base + (uint32_t)kb * 2u * b_tile : sa + 2 * A_TILE_BYTES;
                    const uint64_t ahi = make_kmajor_sw128_desc(sa), alo = make_kmajor_sw128_desc(sa + A_TILE_BYTES);
                    const uint64_t bhi = make_kmajor_sw128_desc(sb), blo = make_kmajor_sw128_desc(sb + b_tile);
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const uint64_t ko = (uint64_t)(k * 2);
                        umma_bf16(td, ahi + ko, bhi + ko, idesc2, (kb | k) ? 1u : 0u);     // A_hi x [B_hi ; B_lo]
                        umma_bf16(tcx, alo + ko, bhi + ko, idesc, 1u);
                    }
                    umma_commit(empty(s));
                    if (++s == p.stages) { s = 0; ph ^= 1u; }
                }
                umma_commit(tm_full(buf));
            }
        }
    } else {
        const int q = warp & 3;
        for (int i = 0; i < my_tiles; ++i) {
            const int buf = i & 1;
            const uint32_t td = tmem_base + (uint32_t)(buf * 2 * p.BN), tcx = td + (uint32_t)p.BN;
            int n0, t0, h0, w0;
            tile_origin(i, n0, t0, h0, w0);
            mbar_wait(tm_full(buf), ((uint32_t)i >> 1) & 1u);
            tc_fence_after();
            tile_epilogue(p, y, accumulate, stats ? stat_smem : nullptr, td, tcx, q, lane, n0, t0, h0, w0, 0);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tm_empty(buf));
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, tmem_cols);
    if (stats) {
        for (int c = threadIdx.x; c < p.BN; c += blockDim.x) {
            if (c < p.Co) {
                atomicAdd(stats + c, (double)stat_smem[c]);
                atomicAdd(stats + p.Co + c, (double)stat_smem[256 + c]);
            }
        }
    }
}

// =============================================================================================
// Halo-patch variant for stride-1 1x3x3 convolutions over 64 channels (layer1 and its dgrads).
//
// The tap-per-box kernels above fetch every activation tile once per filter tap: 9 x 32 KB per 128
// outputs, which saturates the L2 -> SM path (~43 B/clk/SM) long before the tensor pipe (measured: layer1
// conv 35 % tensor-active, 77 GB/s/SM of TMA traffic).  Here ONE TMA box per tile brings the whole input
// patch -- `bhr` image rows x (W + 2) pixels, halo columns / rows zero-filled by the TMA -- into shared
// memory, and the nine taps are nine UMMA descriptors into that same patch:
//
//   * a tile is 128 consecutive positions of the frame in "padded-pitch" order f = h * PW + w, PW = W + 2
//     (positions with w >= W are computed and dropped: 2 / PW of the MMA work);
//   * the patch row of output f under tap (dh, dw) is (f - f0) + rowoff + dh * PW + dw, a CONSTANT row
//     shift per tap, so the K-major SW128 descriptor just starts `shift * 128` bytes further in (the
//     128B swizzle is a function of the absolute shared-memory address, which TMA and UMMA share; the
//     descriptor's base-offset field stays 0 -- measured: bit-exact against the tap-per-box kernel);
//   * weights stream through a ring of per-tap [B_hi ; B_lo] tiles; the two are adjacent, so
//     A_hi x [B_hi ; B_lo] is ONE N = 2*BN instruction writing the main and the cross-term accumulator.
// =============================================================================================
struct HaloParams {
    int PW, bhr;               // padded row pitch (W + 2) and image rows per patch box
    int H, W, T;               // frame extents, frames per clip
    int tiles_per_frame, total_tiles;
    int patch_bytes;           // one plane of the patch, rounded up to 1024
    int bstages;               // weight ring depth
    int BN, Co, Ksrc;
    int shift[9], kcol[9];     // per tap: patch row shift, column of the packed filter matrix
    BnRed red;                 // optional fused reduction (see BnRed)
    long long out_sn, out_st, out_sh, out_sw, out_base;
};

constexpr int HALO_BN = 64, HALO_STAGES = 6;

// MMAs of one tile: 9 taps x 4 k-steps x {A_hi x [B_hi ; B_lo] (N = 128), A_lo x B_hi (N = 64)}.  S0 = ring stage of
// tap 0.  Descriptors are built from their low words (start address >> 4 | LBO field); the high word of a
// K-major SWIZZLE_128B descriptor (SBO = 1024 B, version 1, layout 2) is the constant 0x40004040.
template <int S0>
__device__ __forceinline__ void halo_issue_tile(uint32_t a_lo, const uint32_t (&sh16)[9], uint32_t patch16, uint32_t ring_lo,
                                                uint32_t td, uint32_t tcx, uint32_t idesc2, uint32_t idesc1,
                                                uint32_t bfull0, uint32_t bempty0, uint32_t& ph, uint32_t tmfull) {
    constexpr uint64_t DHI = 0x40004040ull << 32;
#pragma unroll
    for (int t = 0; t < 9; ++t) {
        const int s = (S0 + t) % HALO_STAGES;
        mbar_wait(bfull0 + 8u * s, ph);
        tc_fence_after();
        const uint32_t ahi = a_lo + sh16[t], alo = ahi + patch16, b = ring_lo + (uint32_t)s * (2u * HALO_BN * 128u / 16u);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            umma_bf16(td, DHI | (uint64_t)(ahi + 2 * k), DHI | (uint64_t)(b + 2 * k), idesc2, (t | k) ? 1u : 0u);
            umma_bf16(tcx, DHI | (uint64_t)(alo + 2 * k), DHI | (uint64_t)(b + 2 * k), idesc1, 1u);
        }
        if (s & 1) umma_commit(bempty0 + 8u * (s >> 1));            // releases stages s-1 and s (triples measured slower)
        if (s == HALO_STAGES - 1) ph ^= 1u;
    }
    umma_commit(tmfull);
}

// RED: the epilogue also reduces the BatchNorm-backward sums of the consumer BN (hp.red, see BnRed) -- a separate
// instantiation, so that the prefetch registers of that path do not cost the plain forward / dgrad kernel anything
template <bool RED>
__global__ void __launch_bounds__(320, 1)
conv_tc_halo_kernel(const __grid_constant__ TcMaps maps, const HaloParams hp, float* __restrict__ y, int accumulate,
                    double* __restrict__ stats) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t b_tile = (uint32_t)hp.BN * 128u;
    // smem: [2 x (patch_hi | patch_lo)] [bstages x (B_hi | B_lo)] [barriers] [BN partials]
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t pbuf = 2u * (uint32_t)hp.patch_bytes;
    const uint32_t ring = base + 2u * pbuf;
    const uint32_t bar_base = ring + (uint32_t)hp.bstages * 2u * b_tile;
    auto p_full = [&](int b) { return bar_base + 8u * b; };
    auto p_empty = [&](int b) { return bar_base + 8u * (2 + b); };
    auto tm_full = [&](int b) { return bar_base + 8u * (4 + b); };
    auto tm_empty = [&](int b) { return bar_base + 8u * (6 + b); };
    auto b_full = [&](int s) { return bar_base + 8u * (8 + s); };
    auto b_empty = [&](int s) { return bar_base + 8u * (8 + hp.bstages + s); };
    const uint32_t tmem_ptr_addr = bar_base + 8u * (8 + 2 * hp.bstages);
    float* stat_smem = reinterpret_cast<float*>(smem_raw + (bar_base + 8u * (9 + 2 * hp.bstages) - smem_u32(smem_raw)));
    if (stats) {
        for (int i = threadIdx.x; i < 512; i += blockDim.x) stat_smem[i] = 0.f;
        if (RED && threadIdx.x < HALO_BN) {
            stat_smem[512 + threadIdx.x] = hp.red.mean[threadIdx.x];
            stat_smem[768 + threadIdx.x] = hp.red.rstd[threadIdx.x];
        }
    }
    const uint32_t tmem_cols = 4u * (uint32_t)hp.BN;
    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.a_hi[0]) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.a_lo[0]) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.b_hi) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.b_lo) : "memory");
        for (int b = 0; b < 2; ++b) {
            mbar_init(p_full(b), 1); mbar_init(p_empty(b), 1);
            mbar_init(tm_full(b), 1); mbar_init(tm_empty(b), 8);
        }
        for (int s = 0; s < hp.bstages; ++s) { mbar_init(b_full(s), 1); mbar_init(b_empty(s), 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) tmem_alloc(tmem_ptr_addr, tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<const uint32_t*>(smem_raw + (tmem_ptr_addr - smem_u32(smem_raw)));
    const int my_tiles = ((int)blockIdx.x < hp.total_tiles) ? (hp.total_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;

    // tile i of this CTA -> frame (n, t), first padded-pitch position f0, first image row of the tile
    auto tile_origin = [&](int i, int& n, int& t, int& f0, int& hrow0) {
        const int tile = (int)blockIdx.x + i * (int)gridDim.x;
        const int frame = tile / hp.tiles_per_frame;
        f0 = (tile - frame * hp.tiles_per_frame) * 128;
        hrow0 = f0 / hp.PW;
        n = frame / hp.T; t = frame - n * hp.T;
    };

    // A tcgen05.commit drains the tensor pipe (measured with scripts/umma_bench.cu: ~155 cycles per commit, against
    // 448 cycles of MMAs per tap), so commits are rationed: weight stages are released in PAIRS (one commit per two
    // taps, counted across tile boundaries), and the end-of-tile commit on tm_full doubles as the "patch buffer is
    // free" signal for the producer.
    if (warp == 0) {
        if (elect_one()) {
            const uint32_t patch_tx = 2u * (uint32_t)(hp.bhr * hp.PW) * 128u;
            int next_patch = 0;                      // next tile whose patch has not been requested yet
            auto try_patch = [&](bool block) {
                const int b = next_patch & 1;
                if (next_patch >= 2) {               // buffer b was last read by tile next_patch - 2
                    const uint32_t par = ((uint32_t)(next_patch - 2) >> 1) & 1u;
                    if (block) {
                        mbar_wait(tm_full(b), par);
                    } else {
                        uint32_t ok;
                        asm volatile(
                            "{\n\t.reg .pred P1;\n\t"
                            "mbarrier.test_wait.parity.shared::cta.b64 P1, [%1], %2;\n\t"
                            "selp.u32 %0, 1, 0, P1;\n\t}" : "=r"(ok) : "r"(tm_full(b)), "r"(par) : "memory");
                        if (!ok) return;
                    }
                }
                int n, t, f0, hrow0;
                tile_origin(next_patch, n, t, f0, hrow0);
                mbar_expect_tx(p_full(b), patch_tx);
                tma_load_5d(&maps.a_hi[0], base + b * pbuf, p_full(b), 0, -1, hrow0 - 1, t, n);
                tma_load_5d(&maps.a_lo[0], base + b * pbuf + hp.patch_bytes, p_full(b), 0, -1, hrow0 - 1, t, n);
                ++next_patch;
            };
            if (my_tiles > 0) try_patch(true);
            int s = 0; uint32_t ph = 0;
            for (int i = 0; i < my_tiles; ++i) {
                for (int tap = 0; tap < 9; ++tap) {
                    // the patch of tile i+1 reuses the buffer of tile i-1: request it as soon as that tile has
                    // retired, without stalling the weight ring (the last tap blocks, so it is never skipped)
                    if (next_patch <= i + 1 && next_patch < my_tiles) try_patch(tap == 8);
                    if ((s & 1) == 0) mbar_wait(b_empty(s >> 1), ph ^ 1u);
                    mbar_expect_tx(b_full(s), 2u * b_tile);
                    const uint32_t sb = ring + (uint32_t)s * 2u * b_tile;
                    tma_load_2d(&maps.b_hi, sb, b_full(s), hp.kcol[tap], 0);
                    tma_load_2d(&maps.b_lo, sb + b_tile, b_full(s), hp.kcol[tap], 0);
                    if (++s == hp.bstages) { s = 0; ph ^= 1u; }
                }
            }
        }
    } else if (warp == 1) {
        if (elect_one()) {
            // D = f32, A = B = bf16, K-major; M = 128; N = 2*BN for A_hi x [B_hi ; B_lo], N = BN for A_lo x B_hi
            const uint32_t idesc_n = (1u << 4) | (1u << 7) | (1u << 10) | ((128u >> 4) << 24);
            const uint32_t idesc2 = idesc_n | ((uint32_t)((2 * hp.BN) >> 3) << 17);
            const uint32_t idesc1 = idesc_n | ((uint32_t)(hp.BN >> 3) << 17);
            // One thread issues every MMA, in order, so its own instruction latency is on the critical path: with
            // ~75 dependent instructions per tap the tensor pipe idled ~45 % of the time (measured; weights, patch
            // traffic, stores and commit count all ruled out).  The per-tile body is therefore fully unrolled with
            // compile-time stage indices (6 stages, 9 taps: the ring phase repeats every two tiles) and 32-bit
            // descriptor arithmetic only.
            uint32_t sh16[9];
#pragma unroll
            for (int t = 0; t < 9; ++t) sh16[t] = (uint32_t)hp.shift[t] * 8u;
            const uint32_t patch16 = (uint32_t)hp.patch_bytes >> 4;
            const uint32_t ring_lo = (ring >> 4) | 0x10000u;
            const uint32_t bfull0 = b_full(0), bempty0 = b_empty(0);
            uint32_t ph = 0;
            for (int i = 0; i < my_tiles; ++i) {
                const int buf = i & 1;
                const uint32_t td = tmem_base + (uint32_t)(buf * 2 * HALO_BN), tcx = td + (uint32_t)HALO_BN;
                int n, t, f0, hrow0;
                tile_origin(i, n, t, f0, hrow0);
                const uint32_t a_lo = ((base + buf * pbuf + (uint32_t)(f0 - hrow0 * hp.PW) * 128u) >> 4) | 0x10000u;
                mbar_wait(tm_empty(buf), (((uint32_t)i >> 1) & 1u) ^ 1u);
                mbar_wait(p_full(buf), ((uint32_t)i >> 1) & 1u);
                if (i & 1) halo_issue_tile<3>(a_lo, sh16, patch16, ring_lo, td, tcx, idesc2, idesc1, bfull0, bempty0, ph, tm_full(buf));
                else       halo_issue_tile<0>(a_lo, sh16, patch16, ring_lo, td, tcx, idesc2, idesc1, bfull0, bempty0, ph, tm_full(buf));
            }
        }
    } else {
        // epilogue: 8 warps = 4 TMEM lane quarters (warp % 4) x two 32-column halves.  Both accumulator loads are
        // in flight together and the TMEM buffer is handed back before the global stores; the BatchNorm partial
        // sums stay in registers (one row per lane) across all tiles of the CTA and are transposed once at the end.
        const int q = warp & 3, c0 = ((warp - 2) >> 2) * 32;
        float rs[32], rq[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) { rs[j] = 0.f; rq[j] = 0.f; }
        for (int i = 0; i < my_tiles; ++i) {
            const int buf = i & 1;
            const uint32_t td = tmem_base + (uint32_t)(buf * 2 * hp.BN), tcx = td + (uint32_t)hp.BN;
            int n, t, f0, hrow0;
            tile_origin(i, n, t, f0, hrow0);
            const int f = f0 + q * 32 + lane;
            const int h = f / hp.PW, w = f - h * hp.PW;
            const bool valid = h < hp.H && w < hp.W;
            const long long row = (long long)n * hp.out_sn + (long long)t * hp.out_st + (long long)h * hp.out_sh +
                                  (long long)w * hp.out_sw + hp.out_base;
            // BatchNorm-backward fusion (BnRed): this lane's row of the consumer BN's input `y` and ReLU mask is requested
            // BEFORE waiting for the accumulators, so the row-per-lane gather (32 separate 128-byte lines per warp
            // instruction) overlaps the tile's MMAs instead of stalling the TMEM drain (measured without the prefetch:
            // +1.28 ms on the layer1 dgrad against the 0.52 ms reduce pass it replaces)
            float4 yv[8];
            uint2 mk[8];
            if (RED) {
                const size_t off = (size_t)(valid ? row : 0) * hp.Co + c0;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    yv[j] = *reinterpret_cast<const float4*>(hp.red.y + off + 4 * j);
                    mk[j] = hp.red.mask_hi ? *reinterpret_cast<const uint2*>(hp.red.mask_hi + off + 4 * j)
                                           : make_uint2(0x3c003c00u, 0x3c003c00u);          // +1.0 (fp16): keep
                }
            }
            mbar_wait(tm_full(buf), ((uint32_t)i >> 1) & 1u);
            tc_fence_after();
            if (RED) {
                // two 16-column halves: keeps the accumulator registers (2 x 16) + the prefetched rows (48) + the running
                // sums (64) inside the register file
                float4* dst = reinterpret_cast<float4*>(y + (valid ? row : 0) * hp.Co + c0);
#pragma unroll
                for (int hf = 0; hf < 2; ++hf) {
                    uint32_t v[16], u[16];
                    tmem_ld16_nowait(td + ((uint32_t)(q * 32) << 16) + (uint32_t)(c0 + 16 * hf), v);
                    tmem_ld16_nowait(tcx + ((uint32_t)(q * 32) << 16) + (uint32_t)(c0 + 16 * hf), u);
                    tmem_ld_wait();
                    if (hf == 1) {
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(tm_empty(buf));
                    }
                    if (valid) {
#pragma unroll
                        for (int jj = 0; jj < 4; ++jj) {
                            const int j = 4 * hf + jj;
                            float4 o = make_float4(__uint_as_float(v[4 * jj]) + __uint_as_float(u[4 * jj]),
                                                   __uint_as_float(v[4 * jj + 1]) + __uint_as_float(u[4 * jj + 1]),
                                                   __uint_as_float(v[4 * jj + 2]) + __uint_as_float(u[4 * jj + 2]),
                                                   __uint_as_float(v[4 * jj + 3]) + __uint_as_float(u[4 * jj + 3]));
                            if (accumulate) { const float4 c = dst[j]; o.x += c.x; o.y += c.y; o.z += c.z; o.w += c.w; }
                            dst[j] = o;
                            // BatchNorm-backward sums of the consumer BN (see BnRed)
                            const float* mr = stat_smem + 512 + c0 + 4 * j;
                            const uint32_t h0 = mk[j].x & 0xffffu, h1 = mk[j].x >> 16, h2 = mk[j].y & 0xffffu, h3 = mk[j].y >> 16;
                            const float g0 = ((h0 & 0x8000u) == 0u && (h0 & 0x7fffu) != 0u) ? o.x : 0.f;
                            const float g1 = ((h1 & 0x8000u) == 0u && (h1 & 0x7fffu) != 0u) ? o.y : 0.f;
                            const float g2 = ((h2 & 0x8000u) == 0u && (h2 & 0x7fffu) != 0u) ? o.z : 0.f;
                            const float g3 = ((h3 & 0x8000u) == 0u && (h3 & 0x7fffu) != 0u) ? o.w : 0.f;
                            rs[4 * j] += g0; rs[4 * j + 1] += g1; rs[4 * j + 2] += g2; rs[4 * j + 3] += g3;
                            rq[4 * j] = fmaf(g0, (yv[j].x - mr[0]) * mr[256], rq[4 * j]);
                            rq[4 * j + 1] = fmaf(g1, (yv[j].y - mr[1]) * mr[257], rq[4 * j + 1]);
                            rq[4 * j + 2] = fmaf(g2, (yv[j].z - mr[2]) * mr[258], rq[4 * j + 2]);
                            rq[4 * j + 3] = fmaf(g3, (yv[j].w - mr[3]) * mr[259], rq[4 * j + 3]);
                        }
                    }
                }
            } else {
                uint32_t v[32], u[32];
                tmem_ld32_nowait(td + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
                tmem_ld32_nowait(tcx + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, u);
                tmem_ld_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(tm_empty(buf));
                float4 o[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    o[j] = make_float4(__uint_as_float(v[4 * j]) + __uint_as_float(u[4 * j]),
                                       __uint_as_float(v[4 * j + 1]) + __uint_as_float(u[4 * j + 1]),
                                       __uint_as_float(v[4 * j + 2]) + __uint_as_float(u[4 * j + 2]),
                                       __uint_as_float(v[4 * j + 3]) + __uint_as_float(u[4 * j + 3]));
                    if (stats && valid) {          // BatchNorm statistics of the conv output (forward: no accumulate)
                        rs[4 * j] += o[j].x; rs[4 * j + 1] += o[j].y; rs[4 * j + 2] += o[j].z; rs[4 * j + 3] += o[j].w;
                        rq[4 * j] += o[j].x * o[j].x; rq[4 * j + 1] += o[j].y * o[j].y;
                        rq[4 * j + 2] += o[j].z * o[j].z; rq[4 * j + 3] += o[j].w * o[j].w;
                    }
                }
                pair_store_rows<4>(o, y, row, valid, lane, hp.Co, c0, accumulate != 0);      // full-sector stores
            }
        }
        if (stats) {
            // lane l ends up with column c0 + l summed over the warp's 32 rows (31-shuffle transposing butterfly)
#pragma unroll
            for (int off = 16; off >= 1; off >>= 1) {
                const bool up = (lane & off) != 0;
#pragma unroll
                for (int i = 0; i < off; ++i) {
                    const float s_send = up ? rs[i] : rs[i + off], s_keep = up ? rs[i + off] : rs[i];
                    const float q_send = up ? rq[i] : rq[i + off], q_keep = up ? rq[i + off] : rq[i];
                    rs[i] = s_keep + __shfl_xor_sync(0xffffffffu, s_send, off);
                    rq[i] = q_keep + __shfl_xor_sync(0xffffffffu, q_send, off);
                }
            }
            atomicAdd(&stat_smem[c0 + lane], rs[0]);
            atomicAdd(&stat_smem[256 + c0 + lane], rq[0]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, tmem_cols);
    if (stats) {
        for (int c = threadIdx.x; c < hp.BN; c += blockDim.x) {
            if (c < hp.Co) {
                atomicAdd(stats + c, (double)stat_smem[c]);
                atomicAdd(stats + hp.Co + c, (double)stat_smem[256 + c]);
            }
        }
    }
}

// =============================================================================================
// Halo-patch wgrad for the same sites (stride-1 1x3x3, 64 -> 64 channels):
//
//   dWp[co][tap][ci] = sum over positions  dY[pos, co] * X[pos (+) tap, ci]
//
// K = positions.  A tile is 128 consecutive padded-pitch positions of one frame (as in conv_tc_halo_kernel); ONE
// box brings the X patch (halo included) and one the dY rows, whose out-of-frame columns w >= W and rows h >= H are
// zero-filled by the TMA, so the dropped positions contribute nothing.  Both operands are MN-major views of those
// boxes (a smem row = one position = 64 channels), and a filter tap is again a constant ROW shift of the X operand:
//   A (M = 128) = X^T of TWO taps: the two 64-channel groups of the operand are `LBO` apart, and LBO is simply the
//                 difference of the two taps' row shifts (9 taps = 4 pairs + 1 single, 5 accumulators of 64 columns);
//   B (N = 64)  = dY^T.
// The tap-per-box wgrad kernel fetches X once per tap and pads M = Co = 64 to 128 with zeros; this one fetches
// X once and has no padding (layer1: 2.4 ms -> see profiles/).  All three split-BF16 products of a tap pair share
// one TMEM accumulator (5 x 64 columns; a main/correction split would need 640), so the chain is flushed into the
// fp32 result with atomics every `chain` tiles to bound the truncating in-TMEM accumulation.
// =============================================================================================
struct WgHaloParams {
    int PW, bhr_x, bhr_y, H, W, T;
    int tiles_per_frame, total_tiles;
    int xpatch_bytes, ypatch_bytes;      // one plane, rounded up to 1024
    int chain;                           // tiles per TMEM accumulation chain
    int shift[9];                        // X row shift per tap
};

__global__ void __launch_bounds__(192, 1)
wgrad_halo_kernel(const __grid_constant__ TcMaps maps, const WgHaloParams hp, float* __restrict__ dwp) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // smem: 2 x [x_hi | x_lo | dy_hi | dy_lo] [barriers]
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bufsz = 2u * (uint32_t)hp.xpatch_bytes + 2u * (uint32_t)hp.ypatch_bytes;
    const uint32_t bar_base = base + 2u * bufsz;
    auto full = [&](int b) { return bar_base + 8u * b; };
    auto empty = [&](int b) { return bar_base + 8u * (2 + b); };
    const uint32_t acc_full = bar_base + 32u, acc_empty = bar_base + 40u, tmem_ptr_addr = bar_base + 48u;
    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.a_hi[0]) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.a_lo[0]) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.b_hi) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.b_lo) : "memory");
        for (int b = 0; b < 2; ++b) { mbar_init(full(b), 1); mbar_init(empty(b), 1); }
        mbar_init(acc_full, 1);
        mbar_init(acc_empty, 4);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) tmem_alloc(tmem_ptr_addr, 512);
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
            const uint32_t tx = 2u * (uint32_t)(hp.bhr_x * hp.PW) * 128u + 2u * (uint32_t)(hp.bhr_y * hp.PW) * 128u;
            for (int i = 0; i < my_tiles; ++i) {
                const int b = i & 1;
                int n, t, f0, hrow0;
                tile_origin(i, n, t, f0, hrow0);
                mbar_wait(empty(b), (((uint32_t)i >> 1) & 1u) ^ 1u);
                mbar_expect_tx(full(b), tx);
                const uint32_t sx = base + b * bufsz, sy = sx + 2u * hp.xpatch_bytes;
                tma_load_5d(&maps.a_hi[0], sx, full(b), 0, -1, hrow0 - 1, t, n);
                tma_load_5d(&maps.a_lo[0], sx + hp.xpatch_bytes, full(b), 0, -1, hrow0 - 1, t, n);
                tma_load_5d(&maps.b_hi, sy, full(b), 0, 0, hrow0, t, n);
                tma_load_5d(&maps.b_lo, sy + hp.ypatch_bytes, full(b), 0, 0, hrow0, t, n);
            }
        }
    } else if (warp == 1) {
        if (elect_one()) {
            // D = f32, A = B = bf16, both MN-major (bits 15, 16), N = 64, M = 128
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((64u >> 3) << 17) | ((128u >> 4) << 24);
            constexpr uint64_t DHI = 0x40004040ull << 32;     // SBO = 1024 B, version 1, SWIZZLE_128B
            // low descriptor word of a tap pair relative to the patch: first tap's row shift | LBO = shift difference
            uint32_t pair_lo[5];
#pragma unroll
            for (int p = 0; p < 5; ++p) {
                const int a = 2 * p, b = (2 * p + 1 < 9) ? 2 * p + 1 : a;
                pair_lo[p] = (uint32_t)hp.shift[a] * 8u + (((uint32_t)(hp.shift[b] - hp.shift[a]) * 8u) << 16);
            }
            const uint32_t xpatch16 = (uint32_t)hp.xpatch_bytes >> 4, ypatch16 = (uint32_t)hp.ypatch_bytes >> 4;
            for (int i = 0; i < my_tiles; ++i) {
                const int b = i & 1;
                const int in_chain = i % hp.chain;
                int n, t, f0, hrow0;
                tile_origin(i, n, t, f0, hrow0);
                const uint32_t rowoff16 = (uint32_t)(f0 - hrow0 * hp.PW) * 8u;
                const uint32_t x16 = ((base + b * bufsz) >> 4) + rowoff16;
                const uint32_t y16 = (((base + b * bufsz + 2u * hp.xpatch_bytes) >> 4) + rowoff16) | (512u << 16);
                if (in_chain == 0 && i > 0) {                   // the previous chain must have been drained
                    mbar_wait(acc_empty, (((uint32_t)(i / hp.chain) - 1u) & 1u));
                }
                mbar_wait(full(b), ((uint32_t)i >> 1) & 1u);
                tc_fence_after();
#pragma unroll
                for (int k = 0; k < 8; ++k) {                   // UMMA_K = 16 positions = 2048 B of rows
                    const uint64_t yhi = DHI | (uint64_t)(y16 + k * 128), ylo = DHI | (uint64_t)(y16 + ypatch16 + k * 128);
#pragma unroll
                    for (int p = 0; p < 5; ++p) {
                        const uint64_t xhi = DHI | (uint64_t)(x16 + pair_lo[p] + k * 128);
                        const uint64_t xlo = DHI | (uint64_t)(x16 + xpatch16 + pair_lo[p] + k * 128);
                        const uint32_t d = tmem_base + (uint32_t)(p * 64);
                        umma_bf16(d, xhi, yhi, idesc, (in_chain | k) ? 1u : 0u);
                        umma_bf16(d, xhi, ylo, idesc, 1u);
                        umma_bf16(d, xlo, yhi, idesc, 1u);
                    }
                }
                umma_commit(empty(b));
                if (in_chain == hp.chain - 1 || i == my_tiles - 1) umma_commit(acc_full);
            }
        }
    } else {
        // drain: TMEM lane = (tap of the pair, ci); column = co
        const int q = warp & 3;
        const int m = q * 32 + lane, g = m >> 6, ci = m & 63;
        for (int c = 0; c < chains; ++c) {
            mbar_wait(acc_full, (uint32_t)c & 1u);
            tc_fence_after();
#pragma unroll 1
            for (int p = 0; p < 5; ++p) {
                const int tap = 2 * p + g;
                uint32_t v[32], u[32];
                tmem_ld32_nowait(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(p * 64), v);
                tmem_ld32_nowait(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(p * 64 + 32), u);
                tmem_ld_wait();
                if (tap < 9) {
                    float* o = dwp + (size_t)tap * 64 + ci;
#pragma unroll
                    for (int co = 0; co < 32; ++co) atomicAdd(o + (size_t)co * 576, __uint_as_float(v[co]));
#pragma unroll
                    for (int co = 0; co < 32; ++co) atomicAdd(o + (size_t)(co + 32) * 576, __uint_as_float(u[co]));
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(acc_empty);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, 512);
}

// =============================================================================================
// wgrad:  dW[co][tap][ci] += sum_{positions} dY[pos, co] * X[pos (+) tap, ci]
//   A = dY^T (M = co, K = positions), B = X^T (N = ci, K = positions): both MN-major over the
//   same [positions x 64 channels] TMA boxes.  grid = (co tiles * ci tiles, taps, splits);
//   partial sums are reduced with fp32 atomics.
// =============================================================================================
__global__ void __launch_bounds__(192, 2)
wgrad_tc_kernel(const __grid_constant__ TcMaps maps, const TcParams p, float* __restrict__ dwp) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const int NG = p.BN * p.tap_group;                        // accumulator width: tap_group taps side by side
    const SmemPlan sp = plan_smem(smem_raw, NG, p.stages);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t tmem_cols = 2u * (NG < 16 ? 16 : NG);      // main + correction accumulators
    // Position boxes may hold fewer than a multiple of 16 rows (UMMA_K): rows the TMA never writes
    // must read as zero, so clear the operand stages once (generic proxy), then hand over to the
    // async proxy (TMA / UMMA).
    if (p.box_rows & 15) {
        uint4* z = reinterpret_cast<uint4*>(smem_raw + (sp.base - smem_u32(smem_raw)));
        const int n16 = (int)((uint32_t)sp.stages * sp.stage_bytes / 16u);
        for (int i = threadIdx.x; i < n16; i += blockDim.x) z[i] = make_uint4(0u, 0u, 0u, 0u);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    const uint32_t tmem_d = setup_common(sp, smem_raw, maps, p.nviews, tmem_cols);
    const uint32_t tmem_c = tmem_d + (uint32_t)NG;

    const int ci_tiles = (p.Ksrc + p.BN - 1) / p.BN;
    const int co0 = (blockIdx.x / ci_tiles) * 128, ci0 = (blockIdx.x % ci_tiles) * p.BN;
    // this CTA's taps: [tap_beg, tap_beg + gt).  They share the dY tile (A operand); their shifted X tiles
    // sit side by side in the B operand, so one MMA of width gt*BN covers all of them.
    const int ntaps = p.tT.count * p.tH.count * p.tW.count;
    const int tap_beg = blockIdx.y * p.tap_group;
    const int gt = (ntaps - tap_beg) < p.tap_group ? (ntaps - tap_beg) : p.tap_group;
    const int total_tiles = p.tiles_w * p.tiles_h * p.tiles_t * p.tiles_n;
    const int kt_beg = blockIdx.z * p.ktiles_per_split;
    int kt_end = kt_beg + p.ktiles_per_split;
    if (kt_end > total_tiles) kt_end = total_tiles;
    const int num_kb = kt_end - kt_beg;
    const int nb_groups = p.BN / 64;              // 64-channel groups per tap in the B operand
    const uint32_t box_bytes = (uint32_t)p.box_rows * 128u;   // one [positions x 64ch] box (<= 8 KB)
    const uint32_t GROUP = 8192u;                 // smem distance between 64-channel groups (64 rows * 128 B)

    if (num_kb > 0) {
        if (warp == 0) {
            if (elect_one()) {
                const uint32_t tx = 2u * 2u * box_bytes + 2u * (uint32_t)(gt * nb_groups) * box_bytes;
                int s = 0; uint32_t ph = 0;
                for (int kb = 0; kb < num_kb; ++kb) {
                    int tile = kt_beg + kb;
                    const int tw = tile % p.tiles_w; tile /= p.tiles_w;
                    const int th = tile % p.tiles_h; tile /= p.tiles_h;
                    const int tt = tile % p.tiles_t; tile /= p.tiles_t;
                    const int w0 = tw * p.bw, h0 = th * p.bh, t0 = tt * p.bt, n0 = tile * p.bn;
                    mbar_wait(sp.empty(s), ph ^ 1u);
                    mbar_expect_tx(sp.full(s), tx);
                    const uint32_t sa = sp.base + s * sp.stage_bytes;
                    // A = dY at the output positions (dense map = maps.b_*, 5-D), two 64-channel groups
                    for (int g = 0; g < 2; ++g) {
                        tma_load_5d(&maps.b_hi, sa + g * GROUP, sp.full(s), co0 + g * 64, w0, h0, t0, n0);
                        tma_load_5d(&maps.b_lo, sa + A_TILE_BYTES + g * GROUP, sp.full(s), co0 + g * 64, w0, h0, t0, n0);
                    }
                    // B = X at the tap-shifted positions (parity view), one block of 64-channel groups per tap
                    for (int j = 0; j < gt; ++j) {
                        const int tap = tap_beg + j;
                        const int iw = tap % p.tW.count, ih = (tap / p.tW.count) % p.tH.count, it = tap / (p.tW.count * p.tH.count);
                        const int view = (p.tT.par[it] * p.sH + p.tH.par[ih]) * p.sW + p.tW.par[iw];
                        const int cw = w0 + p.tW.off[iw], chh = h0 + p.tH.off[ih], ct = t0 + p.tT.off[it];
                        for (int g = 0; g < nb_groups; ++g) {
                            const uint32_t off = (uint32_t)(j * nb_groups + g) * GROUP;
                            tma_load_5d(&maps.a_hi[view], sa + 2 * A_TILE_BYTES + off, sp.full(s), ci0 + g * 64, cw, chh, ct, n0);
                            tma_load_5d(&maps.a_lo[view], sa + 2 * A_TILE_BYTES + sp.b_tile_bytes + off, sp.full(s),
                                        ci0 + g * 64, cw, chh, ct, n0);
                        }
                    }
                    if (++s == p.stages) { s = 0; ph ^= 1u; }
                }
            }
        } else if (warp == 1) {
            if (elect_one()) {
                // D=f32, A=B=bf16, both MN-major (bits 15,16), N = gt*BN, M = 128
                const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) |
                                       ((uint32_t)((gt * p.BN) >> 3) << 17) | ((128u >> 4) << 24);
                const int ksteps = (p.box_rows + 15) / 16;  // UMMA_K = 16 positions; rows past the box are zero
                int s = 0; uint32_t ph = 0;
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(sp.full(s), ph);
                    tc_fence_after();
                    const uint32_t sa = sp.base + s * sp.stage_bytes;
                    const uint64_t ahi = make_mnmajor_sw128_desc(sa, GROUP), alo = make_mnmajor_sw128_desc(sa + A_TILE_BYTES, GROUP);
                    const uint64_t bhi = make_mnmajor_sw128_desc(sa + 2 * A_TILE_BYTES, GROUP);
                    const uint64_t blo = make_mnmajor_sw128_desc(sa + 2 * A_TILE_BYTES + sp.b_tile_bytes, GROUP);
                    for (int k = 0; k < ksteps; ++k) {          // 16 positions = 2 swizzle atoms = 2048 B
                        const uint64_t ko = (uint64_t)(k * (2048 >> 4));
                        umma_bf16(tmem_d, ahi + ko, bhi + ko, idesc, (kb | k) ? 1u : 0u);
                        umma_bf16(tmem_c, ahi + ko, blo + ko, idesc, (kb | k) ? 1u : 0u);
                        umma_bf16(tmem_c, alo + ko, bhi + ko, idesc, 1u);
                    }
                    umma_commit(sp.empty(s));
                    if (++s == p.stages) { s = 0; ph ^= 1u; }
                }
                umma_commit(sp.tmem_full());
            }
        } else {
            const int q = warp & 3;
            const int co = co0 + q * 32 + lane;             // D row = output channel
            mbar_wait(sp.tmem_full(), 0);
            tc_fence_after();
            for (int c0 = 0; c0 < gt * p.BN; c0 += 32) {
                const int j = c0 / p.BN, cin = c0 - j * p.BN;
                const int tap = tap_beg + j;
                const int iw = tap % p.tW.count, ih = (tap / p.tW.count) % p.tH.count, it = tap / (p.tW.count * p.tH.count);
                const int tap_full = (p.tT.k[it] * p.kH + p.tH.k[ih]) * p.kW + p.tW.k[iw];
                uint32_t v[32], u[32];
                tmem_ld32(tmem_d + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
                tmem_ld32(tmem_c + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, u);
#pragma unroll
                for (int jj = 0; jj < 32; ++jj) v[jj] = __float_as_uint(__uint_as_float(v[jj]) + __uint_as_float(u[jj]));
                if (co >= p.Co) continue;
                float* orow = dwp + ((size_t)co * p.taps_full + tap_full) * p.Ksrc + ci0 + cin;
#pragma unroll
                for (int jj = 0; jj < 32; ++jj)
                    if (ci0 + cin + jj < p.Ksrc) atomicAdd(orow + jj, __uint_as_float(v[jj]));
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_d, tmem_cols);
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
PFN_cuTensorMapEncodeTiled_v12000 get_encode() {
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

// bf16 tensor, dims[0] contiguous, arbitrary byte strides for the outer dims
int make_map(CUtensorMap* m, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
             const uint32_t* box) {
    auto enc = get_encode();
    DPC_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled entry point not available");
    cuuint64_t gd[5];
    cuuint64_t gs[4];
    cuuint32_t bx[5], es[5];
    for (int i = 0; i < rank; ++i) { gd[i] = dims[i]; bx[i] = box[i]; es[i] = 1; }
    for (int i = 0; i + 1 < rank; ++i) gs[i] = strides_bytes[i];
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), gd, gs, bx, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    DPC_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed (%d) rank %d dims %llu %llu %llu box %u %u %u", (int)r,
                rank, (unsigned long long)dims[0], (unsigned long long)dims[1], (unsigned long long)(rank > 2 ? dims[2] : 0),
                box[0], box[1], rank > 2 ? box[2] : 0);
    return DPC_OK;
}

// Every tile costs a full MMA of `rows_max` rows, so minimise the tile count; ties -> longer rows.
void choose_box(int W, int H, int T, int NB, int rows_max, int& bw, int& bh, int& bt, int& bn) {
    auto cdiv = [](long long a, long long b) { return (a + b - 1) / b; };
    long long best = -1;
    bw = bh = bt = bn = 1;
    for (int w = (W < rows_max ? W : rows_max); w >= 1; --w) {
        const int rh = rows_max / w;
        for (int h = (H < rh ? H : rh); h >= 1; --h) {
            const int rt = rh / h;
            for (int t = (T < rt ? T : rt); t >= 1; --t) {
                const int rn = rt / t;
                const int n = NB < rn ? NB : rn;
                long long cost = cdiv(W, w) * cdiv(H, h) * cdiv(T, t) * cdiv(NB, n);
                if (best < 0 || cost < best) { best = cost; bw = w; bh = h; bt = t; bn = n; }
            }
        }
    }
}

struct TcLaunch {
    TcParams p;
    TcMaps maps;
    dim3 grid;
    size_t smem;
};

// 5-D map over (a parity view of) a channels-last bf16 tensor [NB,T,H,W,C]
int make_act_map(CUtensorMap* m, const void* base, int C, int W, int H, int T, int NB, long long sw, long long sh,
                 long long st, long long sn, const uint32_t* box) {
    const uint64_t d[5] = {(uint64_t)C, (uint64_t)W, (uint64_t)H, (uint64_t)T, (uint64_t)NB};
    const uint64_t s[4] = {(uint64_t)sw * 2, (uint64_t)sh * 2, (uint64_t)st * 2, (uint64_t)sn * 2};
    return make_map(m, base, 5, d, s, box);
}

void set_stages(TcLaunch& L, int num_kb) {
    TcParams& p = L.p;
    const int stage_bytes = 2 * A_TILE_BYTES + 2 * p.BN * 128;
    int stages = (200 * 1024) / stage_bytes;
    // Narrow tiles (BN <= 64: layer1 / its gradients) have only ~9 k-blocks per tile, so the tile prologue and
    // epilogue dominate a one-CTA-per-SM schedule.  Two co-resident CTAs per SM (2 stages of 48 KB each)
    // let one CTA's epilogue overlap the other's MMA main loop.
    if (p.BN <= 64) stages = 2;
    if (stages > 6) stages = 6;
    if (stages > num_kb) stages = num_kb;
    if (stages < 1) stages = 1;
    p.stages = stages;
    L.smem = (size_t)stages * stage_bytes + 8 * (2 * stages + 2) + 4096 /* BN partials, mean, rstd */ + 1024 /* alignment */;
}

int pick_bn(int C) { return C >= 256 ? 256 : (C >= 128 ? 128 : (C >= 64 ? 64 : 32)); }

// forward tap tables for one dimension: input coord = s*o + k - p = s*(o + off) + par
void fwd_taps(TapDim& d, int K, int s, int pad) {
    d.count = (int8_t)K;
    for (int k = 0; k < K; ++k) {
        int par = ((k - pad) % s + s) % s;
        d.k[k] = (int8_t)k;
        d.par[k] = (int8_t)par;
        d.off[k] = (int8_t)((k - pad - par) / s);
    }
}
// dgrad tap tables for output parity class `cls`: taps k = (cls+pad) mod s, +s, ...; dy coord = j + (cls+pad-k)/s
void dgrad_taps(TapDim& d, int K, int s, int pad, int cls) {
    int n = 0;
    for (int k = (cls + pad) % s; k < K; k += s) {
        d.k[n] = (int8_t)k;
        d.par[n] = 0;
        d.off[n] = (int8_t)((cls + pad - k) / s);
        ++n;
    }
    d.count = (int8_t)n;
}

int check_geom(const dpc_conv_geom* g, const char* who) {
    DPC_REQUIRE(g != nullptr, "%s: null geometry", who);
    DPC_REQUIRE(g->kT >= 1 && g->kT <= 3 && g->kH >= 1 && g->kH <= 3 && g->kW >= 1 && g->kW <= 3,
                "%s: filter extents must be 1..3", who);
    DPC_REQUIRE(g->sT >= 1 && g->sT <= 2 && g->sH >= 1 && g->sH <= 2 && g->sW >= 1 && g->sW <= 2,
                "%s: strides must be 1 or 2", who);
    DPC_REQUIRE(g->Ci % 64 == 0 && g->Co % 64 == 0, "%s: Ci (%d) and Co (%d) must be multiples of 64", who, g->Ci, g->Co);
    DPC_REQUIRE((g->Ti + 2 * g->pT - g->kT) / g->sT + 1 == g->To && (g->Hi + 2 * g->pH - g->kH) / g->sH + 1 == g->Ho &&
                    (g->Wi + 2 * g->pW - g->kW) / g->sW + 1 == g->Wo,
                "%s: output extent does not match the conv arithmetic", who);
    return DPC_OK;
}

// parity views of x [NB,Ti,Hi,Wi,Ci] for a conv with strides (sT,sH,sW)
int make_parity_views(TcMaps& maps, const dpc_conv_geom* g, const void* x_hi, const void* x_lo, const uint32_t* box) {
    const long long sw = g->Ci, sh = (long long)g->Wi * sw, st = (long long)g->Hi * sh, sn = (long long)g->Ti * st;
    for (int pt = 0; pt < g->sT; ++pt)
        for (int ph = 0; ph < g->sH; ++ph)
            for (int pw = 0; pw < g->sW; ++pw) {
                const int v = (pt * g->sH + ph) * g->sW + pw;
                const int Tv = (g->Ti - pt + g->sT - 1) / g->sT, Hv = (g->Hi - ph + g->sH - 1) / g->sH,
                          Wv = (g->Wi - pw + g->sW - 1) / g->sW;
                DPC_REQUIRE(Tv > 0 && Hv > 0 && Wv > 0, "parity view is empty");
                const long long off = (pt * st + ph * sh + pw * sw) * 2;   // bytes
                if (int rc = make_act_map(&maps.a_hi[v], (const char*)x_hi + off, g->Ci, Wv, Hv, Tv, g->NB, sw * g->sW,
                                          sh * g->sH, st * g->sT, sn, box))
                    return rc;
                if (int rc = make_act_map(&maps.a_lo[v], (const char*)x_lo + off, g->Ci, Wv, Hv, Tv, g->NB, sw * g->sW,
                                          sh * g->sH, st * g->sT, sn, box))
                    return rc;
            }
    return DPC_OK;
}

// Halo-patch schedule (conv_tc_halo_kernel) for a stride-1 1x3x3 conv over a 64-channel tensor `src`
// [NB,T,H,W,64] (x for the forward, dy for the dgrad) with 64 output channels.  `tH` / `tW` are the tap tables of
// the caller (forward or flipped), `wp_*` the matching packed filter planes [64][9][64].
// Returns 1 if it launched, 0 if the shape is not eligible, < 0 on error.
int try_conv_halo(const void* src_hi, const void* src_lo, const void* wp_hi, const void* wp_lo, int NB, int T, int H,
                  int W, const TapDim& tH, const TapDim& tW, float* y, int accumulate, double* stats, cudaStream_t st,
                  const BnRed* red = nullptr) {
    if (const char* e = getenv("DPC_TC_HALO")) if (!atoi(e)) return 0;
    HaloParams hp;
    memset(&hp, 0, sizeof(hp));
    hp.PW = W + 2;
    hp.bhr = 3 + (129 + hp.PW - 1) / hp.PW;
    if (hp.PW > 256 || hp.bhr > 256 || tH.count != 3 || tW.count != 3) return 0;
    hp.H = H; hp.W = W; hp.T = T;
    hp.tiles_per_frame = (H * hp.PW + 127) / 128;
    const long long total = (long long)NB * T * hp.tiles_per_frame;
    // padded-pitch tiles waste the halo columns and the tail of the last tile of a frame
    if ((double)H * W < 0.6 * 128.0 * hp.tiles_per_frame || total >= (1ll << 31)) return 0;
    hp.total_tiles = (int)total;
    hp.patch_bytes = ((hp.bhr * hp.PW * 128 + 1023) / 1024) * 1024;
    hp.BN = 64; hp.Co = 64; hp.Ksrc = 64;
    const size_t b_stage = 2 * 64 * 128, budget = 225 * 1024;
    const size_t fixed = 4 * (size_t)hp.patch_bytes + 1024 + 4096 + 256;
    if (fixed + 2 * b_stage > budget) return 0;
    if ((int)((budget - fixed) / b_stage) < HALO_STAGES) return 0;   // the kernel is written for a 6-stage weight ring
    hp.bstages = HALO_STAGES;
    for (int ih = 0; ih < 3; ++ih)
        for (int iw = 0; iw < 3; ++iw) {
            if (tH.off[ih] < -1 || tH.off[ih] > 1 || tW.off[iw] < -1 || tW.off[iw] > 1) return 0;
            hp.shift[ih * 3 + iw] = (tH.off[ih] + 1) * hp.PW + (tW.off[iw] + 1);
            hp.kcol[ih * 3 + iw] = (tH.k[ih] * 3 + tW.k[iw]) * 64;
        }
    hp.out_sw = 1; hp.out_sh = W; hp.out_st = (long long)H * W; hp.out_sn = (long long)T * H * W; hp.out_base = 0;
    if (red) hp.red = *red;
    TcMaps maps;
    const uint32_t box[5] = {64, (uint32_t)hp.PW, (uint32_t)hp.bhr, 1, 1};
    const long long sw = 64, sh = (long long)W * sw, sT = (long long)H * sh, sn = (long long)T * sT;
    if (make_act_map(&maps.a_hi[0], src_hi, 64, W, H, T, NB, sw, sh, sT, sn, box)) return -1;
    if (make_act_map(&maps.a_lo[0], src_lo, 64, W, H, T, NB, sw, sh, sT, sn, box)) return -1;
    const uint64_t bd[2] = {(uint64_t)9 * 64, 64};
    const uint64_t bs[1] = {(uint64_t)9 * 64 * 2};
    const uint32_t bb[2] = {64, 64};
    if (make_map(&maps.b_hi, wp_hi, 2, bd, bs, bb)) return -1;
    if (make_map(&maps.b_lo, wp_lo, 2, bd, bs, bb)) return -1;
    const size_t smem = 4 * (size_t)hp.patch_bytes + hp.bstages * b_stage + 8 * (9 + 2 * hp.bstages) + 4096 + 1024;
    auto kern = (red && red->y) ? conv_tc_halo_kernel<true> : conv_tc_halo_kernel<false>;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
        dpc_set_error("conv_tc_halo_kernel: cannot reserve %zu bytes of shared memory", smem);
        return -1;
    }
    const int sms = dpc_num_sms();
    const int grid = hp.total_tiles < sms ? hp.total_tiles : sms;
    kern<<<grid, 320, smem, st>>>(maps, hp, y, accumulate, stats);
    dpc_count_launch(1);
    if (cudaError_t e = cudaGetLastError()) {
        dpc_set_error("conv_tc_halo_kernel launch: %s", cudaGetErrorString(e));
        return -1;
    }
    return 1;
}

// Halo-patch wgrad (wgrad_halo_kernel) for a stride-1 1x3x3 64 -> 64 site; accumulates into the zeroed `dwp`.
// Returns 1 if it launched, 0 if the shape is not eligible, < 0 on error.
int try_wgrad_halo(const void* x_hi, const void* x_lo, const void* dy_hi, const void* dy_lo, int NB, int T, int H, int W,
                   float* dwp, cudaStream_t st) {
    if (const char* e = getenv("DPC_TC_HALO")) if (!atoi(e)) return 0;
    WgHaloParams hp;
    memset(&hp, 0, sizeof(hp));
    hp.PW = W + 2;
    hp.bhr_x = 3 + (129 + hp.PW - 1) / hp.PW;
    hp.bhr_y = (127 + 2 * hp.PW - 1) / hp.PW;                  // rows [rowoff, rowoff + 128), rowoff < PW
    if (hp.PW > 256 || hp.bhr_x > 256) return 0;
    hp.H = H; hp.W = W; hp.T = T;
    hp.tiles_per_frame = (H * hp.PW + 127) / 128;
    const long long total = (long long)NB * T * hp.tiles_per_frame;
    if ((double)H * W < 0.6 * 128.0 * hp.tiles_per_frame || total >= (1ll << 31)) return 0;
    hp.total_tiles = (int)total;
    hp.xpatch_bytes = ((hp.bhr_x * hp.PW * 128 + 1023) / 1024) * 1024;
    hp.ypatch_bytes = ((hp.bhr_y * hp.PW * 128 + 1023) / 1024) * 1024;
    const size_t smem = 2 * (2 * (size_t)hp.xpatch_bytes + 2 * (size_t)hp.ypatch_bytes) + 64 + 1024;
    if (smem > 227 * 1024) return 0;
    hp.chain = 64;              // 64 tiles x 8 k-steps x 3 products = 1536 truncating accumulations per chain
    for (int kh = 0; kh < 3; ++kh)
        for (int kw = 0; kw < 3; ++kw) hp.shift[kh * 3 + kw] = kh * hp.PW + kw;     // x(h + kh - 1, w + kw - 1)
    TcMaps maps;
    const long long sw = 64, sh = (long long)W * sw, sT = (long long)H * sh, sn = (long long)T * sT;
    const uint32_t bx[5] = {64, (uint32_t)hp.PW, (uint32_t)hp.bhr_x, 1, 1};
    const uint32_t by[5] = {64, (uint32_t)hp.PW, (uint32_t)hp.bhr_y, 1, 1};
    if (make_act_map(&maps.a_hi[0], x_hi, 64, W, H, T, NB, sw, sh, sT, sn, bx)) return -1;
    if (make_act_map(&maps.a_lo[0], x_lo, 64, W, H, T, NB, sw, sh, sT, sn, bx)) return -1;
    if (make_act_map(&maps.b_hi, dy_hi, 64, W, H, T, NB, sw, sh, sT, sn, by)) return -1;
    if (make_act_map(&maps.b_lo, dy_lo, 64, W, H, T, NB, sw, sh, sT, sn, by)) return -1;
    if (cudaFuncSetAttribute(wgrad_halo_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
        dpc_set_error("wgrad_halo_kernel: cannot reserve %zu bytes of shared memory", smem);
        return -1;
    }
    const int sms = dpc_num_sms();
    const int grid = hp.total_tiles < sms ? hp.total_tiles : sms;
    wgrad_halo_kernel<<<grid, 192, smem, st>>>(maps, hp, dwp);
    dpc_count_launch(1);
    if (cudaError_t e = cudaGetLastError()) {
        dpc_set_error("wgrad_halo_kernel launch: %s", cudaGetErrorString(e));
        return -1;
    }
    return 1;
}

bool halo_eligible(const dpc_conv_geom* g) {
    return g->kT == 1 && g->kH == 3 && g->kW == 3 && g->sT == 1 && g->sH == 1 && g->sW == 1 && g->pT == 0 && g->pH == 1 &&
           g->pW == 1 && g->Ci == 64 && g->Co == 64;
}

int launch_conv(TcLaunch& L, float* y, int accumulate, cudaStream_t st, double* stats = nullptr) {
    const TcParams& p = L.p;
    const int total_tiles = (int)L.grid.x;
    const int sms = dpc_num_sms();
    if (p.BN <= 128 && L.grid.y == 1 && p.Co <= p.BN && total_tiles >= 2 * sms) {
        // persistent schedule: double-buffered TMEM, resident weights when they fit
        const int num_kb = p.tT.count * p.tH.count * p.tW.count * p.cchunks;
        const size_t b_tile = (size_t)p.BN * 128;
        const size_t w_bytes = (size_t)num_kb * 2 * b_tile;
        const size_t budget = 218 * 1024;
        // resident weights only pay off if >= 4 activation stages still fit (measured: with 2 stages the TMA
        // latency is exposed and streaming the weights through a deeper ring is faster)
        int resident = (w_bytes + 4 * (2 * A_TILE_BYTES) + 4096 <= budget) ? 1 : 0;
        if (const char* e = getenv("DPC_TC_RESIDENT")) resident = resident && atoi(e);     // tuning knob
        const size_t stage_bytes = 2 * A_TILE_BYTES + (resident ? 0 : 2 * b_tile);
        int stages = (int)((budget - 4096 - (resident ? w_bytes : 0)) / stage_bytes);
        if (stages > 6) stages = 6;
        if (stages >= 2) {
            TcParams pp = p;
            pp.stages = stages;
            const size_t smem = (resident ? w_bytes : 0) + (size_t)stages * stage_bytes + 8 * (2 * stages + 6) + 4096 + 1024;
            DPC_CUDA(cudaFuncSetAttribute(conv_tc_persist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            const int grid = total_tiles < sms ? total_tiles : sms;
            conv_tc_persist_kernel<<<grid, 192, smem, st>>>(L.maps, pp, y, accumulate, stats, resident, total_tiles);
            DPC_LAUNCH_CHECK();
            return DPC_OK;
        }
    }
    DPC_CUDA(cudaFuncSetAttribute(conv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.smem));
    conv_tc_kernel<<<L.grid, 192, L.smem, st>>>(L.maps, L.p, y, accumulate, stats);
    DPC_LAUNCH_CHECK();
    return DPC_OK;
}

void set_tile_grid(TcParams& p, int NB, int T, int H, int W, int rows_max) {
    choose_box(W, H, T, NB, rows_max, p.bw, p.bh, p.bt, p.bn);
    p.box_rows = p.bw * p.bh * p.bt * p.bn;
    p.tiles_w = (W + p.bw - 1) / p.bw; p.tiles_h = (H + p.bh - 1) / p.bh; p.tiles_t = (T + p.bt - 1) / p.bt;
    p.tiles_n = (NB + p.bn - 1) / p.bn;
    p.NB = NB; p.To = T; p.Ho = H; p.Wo = W;
}

template <bool F16>
__global__ void split_bf16_kernel(const float4* __restrict__ src, uint2* __restrict__ hi, uint2* __restrict__ lo,
                                  long long n4) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        float4 v = src[i];
        float f[4] = {v.x, v.y, v.z, v.w};
        unsigned short h[4], l[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (F16) {
                const __half bh = __float2half_rn(f[j]);
                const __half bl = __float2half_rn(f[j] - __half2float(bh));
                h[j] = __half_as_ushort(bh);
                l[j] = __half_as_ushort(bl);
            } else {
                __nv_bfloat16 bh = __float2bfloat16_rn(f[j]);
                __nv_bfloat16 bl = __float2bfloat16_rn(f[j] - __bfloat162float(bh));
                h[j] = __bfloat16_as_ushort(bh);
                l[j] = __bfloat16_as_ushort(bl);
            }
        }
        hi[i] = make_uint2((uint32_t)h[0] | ((uint32_t)h[1] << 16), (uint32_t)h[2] | ((uint32_t)h[3] << 16));
        lo[i] = make_uint2((uint32_t)l[0] | ((uint32_t)l[1] << 16), (uint32_t)l[2] | ((uint32_t)l[3] << 16));
    }
}

// w [Co][Ci][taps] fp32 -> forward planes [Co][tap][Ci] and dgrad planes [Ci][tap][Co]
__global__ void pack_weight_bf16_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ fh,
                                        __nv_bfloat16* __restrict__ fl, __nv_bfloat16* __restrict__ dh,
                                        __nv_bfloat16* __restrict__ dl, int Co, int Ci, int taps) {
    long long total = (long long)Co * Ci * taps;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        int ci = (int)(i % Ci);
        int tap = (int)((i / Ci) % taps);
        int co = (int)(i / ((long long)Ci * taps));
        float v = w[((size_t)co * Ci + ci) * taps + tap];
        __nv_bfloat16 h = __float2bfloat16_rn(v);
        __nv_bfloat16 l = __float2bfloat16_rn(v - __bfloat162float(h));
        if (fh) { fh[i] = h; fl[i] = l; }
        if (dh) {
            size_t j = ((size_t)ci * taps + tap) * Co + co;
            dh[j] = h; dl[j] = l;
        }
    }
}

// dwp [Co][tap][Ci] -> dw [Co][Ci][tap]
__global__ void unpack_wgrad_ctc_kernel(const float* __restrict__ dwp, float* __restrict__ dw, int Co, int Ci, int taps) {
    long long total = (long long)Co * Ci * taps;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        int ci = (int)(i % Ci);
        int tap = (int)((i / Ci) % taps);
        int co = (int)(i / ((long long)Ci * taps));
        dw[((size_t)co * Ci + ci) * taps + tap] = dwp[i];
    }
}

}  // namespace

extern "C" int dpc_split_bf16(const float* src, void* hi, void* lo, int64_t n, void* stream) {
    DPC_REQUIRE(src && hi && lo && n > 0 && n % 4 == 0, "dpc_split_bf16: bad args (n must be a multiple of 4)");
    long long n4 = n / 4;
    long long blocks = (n4 + 255) / 256;
    long long cap = (long long)dpc_num_sms() * 16;
    split_bf16_kernel<false><<<(int)(blocks < cap ? blocks : cap), 256, 0, as_stream(stream)>>>(
        (const float4*)src, (uint2*)hi, (uint2*)lo, n4);
    DPC_LAUNCH_CHECK();
    return DPC_OK;
}

// the same split into fp16 pairs (hi = fp16(x), lo = fp16(x - hi): ~22 mantissa bits): forward-value operands
extern "C" int dpc_split_f16(const float* src, void* hi, void* lo, int64_t n, void* stream) {
    DPC_REQUIRE(src && hi && lo && n > 0 && n % 4 == 0, "dpc_split_f16: bad args (n must be a multiple of 4)");
    long long n4 = n / 4;
    long long blocks = (n4 + 255) / 256;
    long long cap = (long long)dpc_num_sms() * 16;
    split_bf16_kernel<true><<<(int)(blocks < cap ? blocks : cap), 256, 0, as_stream(stream)>>>(
        (const float4*)src, (uint2*)hi, (uint2*)lo, n4);
    DPC_LAUNCH_CHECK();
    return DPC_OK;
}

extern "C" int dpc_pack_conv_weight_bf16(const float* w, void* wf_hi, void* wf_lo, void* wd_hi, void* wd_lo,
                                         int Co, int Ci, int taps, void* stream) {
    DPC_REQUIRE(w && ((wf_hi && wf_lo) || (wd_hi && wd_lo)) && Co > 0 && Ci > 0 && taps > 0,
                "dpc_pack_conv_weight_bf16: bad args");
    long long total = (long long)Co * Ci * taps;
    long long blocks = (total + 255) / 256;
    long long cap = (long long)dpc_num_sms() * 8;
    pack_weight_bf16_kernel<<<(int)(blocks < cap ? blocks : cap), 256, 0, as_stream(stream)>>>(
        w, (__nv_bfloat16*)wf_hi, (__nv_bfloat16*)wf_lo, (__nv_bfloat16*)wd_hi, (__nv_bfloat16*)wd_lo, Co, Ci, taps);
    DPC_LAUNCH_CHECK();
    return DPC_OK;
}

// C[M,N] (fp32, ldc = N) (+)= A[M,K] * B[N,K]^T from split 16-bit planes (K % 64 == 0); f16 != 0: BOTH operands are fp16
// pairs (dpc_split_f16; forward values of O(1) magnitude: ~22 mantissa bits), else bf16 pairs (dpc_split_bf16)
extern "C" int dpc_gemm_nt_split_tc(int M, int N, int K, const void* a_hi, const void* a_lo, const void* b_hi,
                                    const void* b_lo, int f16, float* C, int accumulate, void* stream) {
    DPC_REQUIRE(M > 0 && N > 0 && K > 0 && K % 64 == 0, "dpc_gemm_nt_split_tc: bad dims %d %d %d", M, N, K);
    DPC_REQUIRE(a_hi && a_lo && b_hi && b_lo && C, "dpc_gemm_nt_split_tc: null pointer");
    TcLaunch L;
    memset(&L.p, 0, sizeof(L.p));
    TcParams& p = L.p;
    // A is a [1,1,1,M,K] "image" convolved with a 1x1x1 bank of N filters
    fwd_taps(p.tT, 1, 1, 0); fwd_taps(p.tH, 1, 1, 0); fwd_taps(p.tW, 1, 1, 0);
    p.kH = p.kW = 1; p.sH = p.sW = 1; p.nviews = 1;
    p.cchunks = K / 64; p.Ksrc = K;
    p.ab_f16 = f16 ? 1 : 0;
    set_tile_grid(p, 1, 1, 1, M, 128);
    p.Co = N; p.BN = pick_bn(N);
    set_stages(L, p.cchunks);
    p.out_sw = 1; p.out_sh = M; p.out_st = M; p.out_sn = M; p.out_base = 0;
    const uint32_t box[5] = {64, (uint32_t)p.bw, 1, 1, 1};
    const long long mk = (long long)M * K;
    if (int rc = make_act_map(&L.maps.a_hi[0], a_hi, K, M, 1, 1, 1, K, mk, mk, mk, box)) return rc;
    if (int rc = make_act_map(&L.maps.a_lo[0], a_lo, K, M, 1, 1, 1, K, mk, mk, mk, box)) return rc;
    const uint64_t bd[2] = {(uint64_t)K, (uint64_t)N};
    const uint64_t bs[1] = {(uint64_t)K * 2};
    const uint32_t bb[2] = {64, (uint32_t)p.BN};
    if (int rc = make_map(&L.maps.b_hi, b_hi, 2, bd, bs, bb)) return rc;
    if (int rc = make_map(&L.maps.b_lo, b_lo, 2, bd, bs, bb)) return rc;
    L.grid = dim3((unsigned)(p.tiles_w), (unsigned)((N + p.BN - 1) / p.BN));
    return launch_conv(L, C, accumulate, as_stream(stream));
}

// forward conv, any stride in {1,2}: y [NB,To,Ho,Wo,Co] = conv(x planes [NB,Ti,Hi,Wi,Ci], wf planes [Co][taps][Ci])
extern "C" int dpc_conv3d_fwd_tc(const dpc_conv_geom* g, const void* x_hi, const void* x_lo, const void* wf_hi,
                                 const void* wf_lo, float* y, double* bn_ws, void* stream) {
    if (int rc = check_geom(g, "dpc_conv3d_fwd_tc")) return rc;
    DPC_REQUIRE(x_hi && x_lo && wf_hi && wf_lo && y, "dpc_conv3d_fwd_tc: null pointer");
    TcLaunch L;
    memset(&L.p, 0, sizeof(L.p));
    TcParams& p = L.p;
    fwd_taps(p.tT, g->kT, g->sT, g->pT); fwd_taps(p.tH, g->kH, g->sH, g->pH); fwd_taps(p.tW, g->kW, g->sW, g->pW);
    p.kH = g->kH; p.kW = g->kW; p.sH = g->sH; p.sW = g->sW; p.nviews = g->sT * g->sH * g->sW;
    p.cchunks = g->Ci / 64; p.Ksrc = g->Ci;
    if (halo_eligible(g)) {
        if (bn_ws) DPC_CUDA(cudaMemsetAsync(bn_ws, 0, sizeof(double) * 2 * g->Co, as_stream(stream)));
        const int r = try_conv_halo(x_hi, x_lo, wf_hi, wf_lo, g->NB, g->To, g->Ho, g->Wo, p.tH, p.tW, y, 0, bn_ws,
                                    as_stream(stream));
        if (r < 0) return DPC_ERR_CUDA;
        if (r > 0) return DPC_OK;
    }
    set_tile_grid(p, g->NB, g->To, g->Ho, g->Wo, 128);
    p.Co = g->Co; p.BN = pick_bn(g->Co);
    const int taps = g->kT * g->kH * g->kW;
    set_stages(L, taps * p.cchunks);
    p.out_sw = 1; p.out_sh = g->Wo; p.out_st = (long long)g->Ho * g->Wo; p.out_sn = (long long)g->To * g->Ho * g->Wo;
    p.out_base = 0;
    const uint32_t box[5] = {64, (uint32_t)p.bw, (uint32_t)p.bh, (uint32_t)p.bt, (uint32_t)p.bn};
    if (int rc = make_parity_views(L.maps, g, x_hi, x_lo, box)) return rc;
    const uint64_t bd[2] = {(uint64_t)taps * g->Ci, (uint64_t)g->Co};
    const uint64_t bs[1] = {(uint64_t)taps * g->Ci * 2};
    const uint32_t bb[2] = {64, (uint32_t)p.BN};
    if (int rc = make_map(&L.maps.b_hi, wf_hi, 2, bd, bs, bb)) return rc;
    if (int rc = make_map(&L.maps.b_lo, wf_lo, 2, bd, bs, bb)) return rc;
    L.grid = dim3((unsigned)(p.tiles_w * p.tiles_h * p.tiles_t * p.tiles_n), (unsigned)((g->Co + p.BN - 1) / p.BN));
    if (bn_ws) DPC_CUDA(cudaMemsetAsync(bn_ws, 0, sizeof(double) * 2 * g->Co, as_stream(stream)));
    return launch_conv(L, y, 0, as_stream(stream), bn_ws);
}

// dgrad, any stride in {1,2}: dx [NB,Ti,Hi,Wi,Ci] (+)= conv^T(dy planes [NB,To,Ho,Wo,Co], wd planes [Ci][taps][Co]).
// One launch per input-parity class; with accumulate == 0 classes without taps are NOT written
// (only possible for 1x1 strided sites, which the caller accumulates into an existing dx).
static int dgrad_tc_impl(const dpc_conv_geom* g, const void* dy_hi, const void* dy_lo, const void* wd_hi,
                         const void* wd_lo, float* dx, int accumulate, const BnRed* red, double* ws, void* stream) {
    if (int rc = check_geom(g, "dpc_conv3d_dgrad_tc")) return rc;
    DPC_REQUIRE(dy_hi && dy_lo && wd_hi && wd_lo && dx, "dpc_conv3d_dgrad_tc: null pointer");
    if (red) {
        DPC_REQUIRE(g->sT == 1 && g->sH == 1 && g->sW == 1, "dpc_conv3d_dgrad_bnred_tc: stride-1 sites only");
        DPC_REQUIRE(red->y && red->mean && red->rstd && ws, "dpc_conv3d_dgrad_bnred_tc: null pointer");
        DPC_CUDA(cudaMemsetAsync(ws, 0, sizeof(double) * 2 * g->Ci, as_stream(stream)));
    }
    const int taps = g->kT * g->kH * g->kW;
    const long long rsw = 1, rsh = g->Wi, rst = (long long)g->Hi * g->Wi, rsn = (long long)g->Ti * g->Hi * g->Wi;
    for (int ct = 0; ct < g->sT; ++ct)
        for (int chh = 0; chh < g->sH; ++chh)
            for (int cw = 0; cw < g->sW; ++cw) {
                TcLaunch L;
                memset(&L.p, 0, sizeof(L.p));
                TcParams& p = L.p;
                dgrad_taps(p.tT, g->kT, g->sT, g->pT, ct);
                dgrad_taps(p.tH, g->kH, g->sH, g->pH, chh);
                dgrad_taps(p.tW, g->kW, g->sW, g->pW, cw);
                const int Tc = (g->Ti - ct + g->sT - 1) / g->sT, Hc = (g->Hi - chh + g->sH - 1) / g->sH,
                          Wc = (g->Wi - cw + g->sW - 1) / g->sW;
                if (Tc <= 0 || Hc <= 0 || Wc <= 0) continue;
                if (p.tT.count == 0 || p.tH.count == 0 || p.tW.count == 0) {
                    DPC_REQUIRE(accumulate, "dpc_conv3d_dgrad_tc: parity class without taps needs accumulate != 0");
                    continue;
                }
                p.kH = g->kH; p.kW = g->kW; p.sH = 1; p.sW = 1; p.nviews = 1;
                p.cchunks = g->Co / 64; p.Ksrc = g->Co;
                if (halo_eligible(g)) {
                    const int r = try_conv_halo(dy_hi, dy_lo, wd_hi, wd_lo, g->NB, g->Ti, g->Hi, g->Wi, p.tH, p.tW, dx,
                                                accumulate, red ? ws : nullptr, as_stream(stream), red);
                    if (r < 0) return DPC_ERR_CUDA;
                    if (r > 0) continue;
                }
                set_tile_grid(p, g->NB, Tc, Hc, Wc, 128);
                p.Co = g->Ci; p.BN = pick_bn(g->Ci);
                set_stages(L, p.tT.count * p.tH.count * p.tW.count * p.cchunks);
                p.out_sw = rsw * g->sW; p.out_sh = rsh * g->sH; p.out_st = rst * g->sT; p.out_sn = rsn;
                p.out_base = ct * rst + chh * rsh + cw * rsw;
                const uint32_t box[5] = {64, (uint32_t)p.bw, (uint32_t)p.bh, (uint32_t)p.bt, (uint32_t)p.bn};
                const long long sw = g->Co, sh = (long long)g->Wo * sw, st = (long long)g->Ho * sh, sn = (long long)g->To * st;
                if (int rc = make_act_map(&L.maps.a_hi[0], dy_hi, g->Co, g->Wo, g->Ho, g->To, g->NB, sw, sh, st, sn, box)) return rc;
                if (int rc = make_act_map(&L.maps.a_lo[0], dy_lo, g->Co, g->Wo, g->Ho, g->To, g->NB, sw, sh, st, sn, box)) return rc;
                const uint64_t bd[2] = {(uint64_t)taps * g->Co, (uint64_t)g->Ci};
                const uint64_t bs[1] = {(uint64_t)taps * g->Co * 2};
                const uint32_t bb[2] = {64, (uint32_t)p.BN};
                if (int rc = make_map(&L.maps.b_hi, wd_hi, 2, bd, bs, bb)) return rc;
                if (int rc = make_map(&L.maps.b_lo, wd_lo, 2, bd, bs, bb)) return rc;
                L.grid = dim3((unsigned)(p.tiles_w * p.tiles_h * p.tiles_t * p.tiles_n), (unsigned)((g->Ci + p.BN - 1) / p.BN));
                if (red) p.red = *red;
                if (int rc = launch_conv(L, dx, accumulate, as_stream(stream), red ? ws : nullptr)) return rc;
            }
    return DPC_OK;
}

extern "C" int dpc_conv3d_dgrad_tc(const dpc_conv_geom* g, const void* dy_hi, const void* dy_lo, const void* wd_hi,
                                   const void* wd_lo, float* dx, int accumulate, void* stream) {
    return dgrad_tc_impl(g, dy_hi, dy_lo, wd_hi, wd_lo, dx, accumulate, nullptr, nullptr, stream);
}

// dgrad of a STRIDE-1 site with the BatchNorm-backward reduction of the consumer BN fused into the epilogue:
// with v = the final dx (after `accumulate`), g = v * [mask_hi > 0] (mask_hi nullable: no ReLU) and
// xhat = (y - mean) * rstd, ws [2*Ci doubles] = sum_rows g | sum_rows g * xhat   (what dpc_bn_bwd's reduce pass
// computes from dx, `out`, `y`); mask_hi / y are [rows, Ci] like dx.
extern "C" int dpc_conv3d_dgrad_bnred_tc(const dpc_conv_geom* g, const void* dy_hi, const void* dy_lo, const void* wd_hi,
                                         const void* wd_lo, float* dx, int accumulate, const void* mask_hi, const float* y,
                                         const float* mean, const float* rstd, double* ws, void* stream) {
    BnRed red;
    red.mask_hi = reinterpret_cast<const __nv_bfloat16*>(mask_hi);
    red.y = y; red.mean = mean; red.rstd = rstd;
    return dgrad_tc_impl(g, dy_hi, dy_lo, wd_hi, wd_lo, dx, accumulate, &red, ws, stream);
}

// wgrad, any stride in {1,2}: dw [Co,Ci,kT,kH,kW] = sum_positions dy (x) x; `dwp` = scratch [Co][taps][Ci] fp32
extern "C" int dpc_conv3d_wgrad_tc(const dpc_conv_geom* g, const void* x_hi, const void* x_lo, const void* dy_hi,
                                   const void* dy_lo, float* dwp, float* dw, void* stream) {
    if (int rc = check_geom(g, "dpc_conv3d_wgrad_tc")) return rc;
    DPC_REQUIRE(x_hi && x_lo && dy_hi && dy_lo && dwp && dw, "dpc_conv3d_wgrad_tc: null pointer");
    cudaStream_t st = as_stream(stream);
    const int taps = g->kT * g->kH * g->kW;
    if (halo_eligible(g)) {
        DPC_CUDA(cudaMemsetAsync(dwp, 0, sizeof(float) * (size_t)g->Co * taps * g->Ci, st));
        const int r = try_wgrad_halo(x_hi, x_lo, dy_hi, dy_lo, g->NB, g->To, g->Ho, g->Wo, dwp, st);
        if (r < 0) return DPC_ERR_CUDA;
        if (r > 0) {
            const long long total = (long long)g->Co * g->Ci * taps;
            const long long blocks = (total + 255) / 256, cap = (long long)dpc_num_sms() * 8;
            unpack_wgrad_ctc_kernel<<<(int)(blocks < cap ? blocks : cap), 256, 0, st>>>(dwp, dw, g->Co, g->Ci, taps);
            DPC_LAUNCH_CHECK();
            return DPC_OK;
        }
    }
    TcLaunch L;
    memset(&L.p, 0, sizeof(L.p));
    TcParams& p = L.p;
    fwd_taps(p.tT, g->kT, g->sT, g->pT); fwd_taps(p.tH, g->kH, g->sH, g->pH); fwd_taps(p.tW, g->kW, g->sW, g->pW);
    p.kH = g->kH; p.kW = g->kW; p.sH = g->sH; p.sW = g->sW; p.nviews = g->sT * g->sH * g->sW;
    p.Ksrc = g->Ci; p.cchunks = g->Ci / 64;
    set_tile_grid(p, g->NB, g->To, g->Ho, g->Wo, 64);           // K-block = up to 64 output positions
    p.Co = g->Co; p.BN = pick_bn(g->Ci);
    p.taps_full = taps;
    // narrow inputs: several taps share one CTA (and one dY tile); their accumulators sit side by side (N <= 256)
    p.tap_group = 256 / p.BN;
    if (p.tap_group > taps) p.tap_group = taps;
    if (p.tap_group < 1) p.tap_group = 1;
    const int tap_groups = (taps + p.tap_group - 1) / p.tap_group;
    const int total_tiles = p.tiles_w * p.tiles_h * p.tiles_t * p.tiles_n;
    const int work = ((g->Co + 127) / 128) * ((g->Ci + p.BN - 1) / p.BN) * tap_groups;
    // K splits: one CTA per SM is resident, so pick the split count that wastes the least of the last wave
    // (time ~ ceil(work*s / SMs) / s), among counts that leave >= 8 position tiles per CTA
    int splits = 1;
    {
        const int sms = dpc_num_sms();
        int s_lo = (sms + work - 1) / work, s_hi = (6 * sms + work - 1) / work;
        if (s_hi > total_tiles / 8) s_hi = total_tiles / 8;
        if (s_lo > s_hi) s_lo = s_hi;
        if (s_lo < 1) s_lo = 1;
        double best = 1e30;
        for (int sp = s_lo; sp <= (s_hi > s_lo ? s_hi : s_lo); ++sp) {
            const long long ctas = (long long)work * sp;
            const double cost = (double)((ctas + sms - 1) / sms) / (double)sp;
            if (cost < best * 0.98) { best = cost; splits = sp; }
        }
    }
    if (splits > total_tiles) splits = total_tiles;
    p.ktiles_per_split = (total_tiles + splits - 1) / splits;
    // bound the in-TMEM accumulation chain (truncating adds): <= 512 position tiles (2048 UMMA steps,
    // ~ -4e-5 relative); longer reductions continue through the fp32 (round-to-nearest) atomics
    int max_kt = 512;                                         // (more, shorter CTAs measured slower: prologue + atomics)
    if (const char* e = getenv("DPC_WGRAD_MAX_KTILES")) if (atoi(e) > 0) max_kt = atoi(e);     // test knob: force the cap
    if (p.ktiles_per_split > max_kt) p.ktiles_per_split = max_kt;
    splits = (total_tiles + p.ktiles_per_split - 1) / p.ktiles_per_split;
    DPC_REQUIRE(splits <= 65535, "dpc_conv3d_wgrad_tc: too many K splits (%d)", splits);
    p.splits = splits;
    {
        const int bn = p.BN;
        p.BN = bn * p.tap_group;               // stage sizing uses the grouped width
        set_stages(L, p.ktiles_per_split);
        p.BN = bn;
    }
    const uint32_t box[5] = {64, (uint32_t)p.bw, (uint32_t)p.bh, (uint32_t)p.bt, (uint32_t)p.bn};
    if (int rc = make_parity_views(L.maps, g, x_hi, x_lo, box)) return rc;
    const long long sw = g->Co, sh = (long long)g->Wo * sw, sT = (long long)g->Ho * sh, sn = (long long)g->To * sT;
    if (int rc = make_act_map(&L.maps.b_hi, dy_hi, g->Co, g->Wo, g->Ho, g->To, g->NB, sw, sh, sT, sn, box)) return rc;
    if (int rc = make_act_map(&L.maps.b_lo, dy_lo, g->Co, g->Wo, g->Ho, g->To, g->NB, sw, sh, sT, sn, box)) return rc;
    DPC_CUDA(cudaMemsetAsync(dwp, 0, sizeof(float) * (size_t)g->Co * taps * g->Ci, st));
    L.grid = dim3((unsigned)(((g->Co + 127) / 128) * ((g->Ci + p.BN - 1) / p.BN)), (unsigned)tap_groups, (unsigned)splits);
    DPC_CUDA(cudaFuncSetAttribute(wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.smem));
    wgrad_tc_kernel<<<L.grid, 192, L.smem, st>>>(L.maps, L.p, dwp);
    DPC_LAUNCH_CHECK();
    long long total = (long long)g->Co * g->Ci * taps;
    long long blocks = (total + 255) / 256;
    long long cap = (long long)dpc_num_sms() * 8;
    unpack_wgrad_ctc_kernel<<<(int)(blocks < cap ? blocks : cap), 256, 0, st>>>(dwp, dw, g->Co, g->Ci, taps);
    DPC_LAUNCH_CHECK();
    return DPC_OK;
}

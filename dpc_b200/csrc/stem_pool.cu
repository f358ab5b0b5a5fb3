// Stem with the conv1 output NEVER stored: conv1 (stem_s2d.cu's stride-1 4x4 conv over space-to-depth planes) fused
// with bn1 statistics + ReLU + MaxPool3d((1,3,3),s(1,2,2),p(0,1,1)) in the forward, and recomputed in the backward.
//
// At BASELINE config 2 the conv1 output is 5.4 GB: round 1 wrote it, re-read it for the pool, re-read it twice in the
// tail backward and wrote / re-read a 5.4 GB gradient on the same grid (>= 27 GB of HBM traffic around a tensor whose
// producer costs ~1 ms of tensor time).  Here:
//
//   forward   stem_pool_fwd_kernel: conv tile in TMEM -> fp32 ring in shared memory (3 tiles: a pooling window spans
//             <= 3 tiles) -> per pooled position the SELECTED raw conv value and its window index.  The selection needs
//             no statistics: relu(bn(.)) is monotone per channel, so max-pool o relu o bn = relu o bn o (max if
//             gamma >= 0 else min) -- the kernel takes max of sign(gamma) * y.  bn1's batch statistics over ALL conv
//             positions come from the same epilogue.  stem_pool_finalize_kernel then normalises the pooled grid (1/4 of
//             the positions) straight into layer1's split-bf16 operand planes and marks ReLU-dead windows in the index.
//   backward  stem_pool_bwd_reduce_kernel: bn1's backward sums on the pooled grid (the gradient of a window lands on its
//             selected position, whose conv value was kept).  stem_pool_bwd_kernel: conv recomputed tile by tile; the
//             epilogue gathers each position's pooled gradient from a TMA-staged window of (dout, index) rows, applies
//             the BatchNorm backward and writes the split-bf16 gradient planes the conv1 wgrad reads.
//
// Frames wider than ~80 conv columns are cut into column bands (one shared-memory pitch per band) so that the ring and
// the input patch fit: 224^2 inputs run as two bands of 56 / 57 columns.
//
// Replaces conv1 / bn1 / relu / maxpool at backbone/resnet_2d3d.py:211-214,260-263 (forward and backward).
#include "tc_common.cuh"

namespace {

constexpr int SP_CH = 16, SP_TAPS = 16, SP_BN = 64, SP_EC = 16;
constexpr int SP_THREADS = 64 + 32 * 4 * (SP_BN / SP_EC);          // producer + MMA warp + 16 epilogue warps
constexpr int SP_EPI = SP_THREADS - 64;
constexpr uint32_t SP_W_BYTES = SP_TAPS * 2 * SP_BN * 32;          // filter bank: per tap [W_hi ; W_lo] x 32 B = 64 KB
constexpr uint64_t SP_DHI = 0xC0004010ull << 32;                   // K-major SWIZZLE_32B descriptor, high word (stem_s2d.cu)
constexpr int RING_PITCH = 272;                                    // 64 fp32 + 16 B pad: conflict-free 16-byte accesses
constexpr int RING_SLOT = 128 * RING_PITCH, RING_SLOTS = 3;
constexpr int SP_MAX_BANDS = 4, SP_NPB_MAX = 4;

struct SpMaps { CUtensorMap x_hi, x_lo, w, dout, didx; };

struct SpGeom {
    int Ho, Wo, T, Hp, Wp;                 // conv1 output frame, frames per block, pooled frame
    int nbands, pitch, bhr;                // column bands per frame, shared-memory row pitch, rows per patch box
    int c0[SP_MAX_BANDS];                  // first conv column of the band
    int cols[SP_MAX_BANDS];                // conv columns computed by the band
    int own0[SP_MAX_BANDS];                // first conv column OWNED by the band (statistics / gradient stores)
    int wp0[SP_MAX_BANDS], npc[SP_MAX_BANDS];      // pooled columns [wp0, wp0 + npc)
    int tiles_per_unit, total_units;       // unit = (frame, band)
    int patch_bytes, npb;
    int shift[SP_TAPS];
    int bc, br, dhalf_bytes, didx_bytes;   // backward: pooled-window box (columns, rows) and its stage sizes
    double inv_n;                          // 1 / (conv positions per channel)
};

struct TileCoord { int n, t, frame, band, i, f0, hrow0; };

__device__ __forceinline__ TileCoord sp_tile(const SpGeom& g, int it) {
    TileCoord c;
    const int j = it / g.tiles_per_unit;
    c.i = it - j * g.tiles_per_unit;
    const int u = (int)blockIdx.x + j * (int)gridDim.x;
    c.frame = u / g.nbands;
    c.band = u - c.frame * g.nbands;
    c.n = c.frame / g.T;
    c.t = c.frame - c.n * g.T;
    c.f0 = c.i * 128;
    c.hrow0 = c.f0 / g.pitch;
    return c;
}

__device__ __forceinline__ void sp_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void sp_epi_sync() { asm volatile("bar.sync 1, %0;" ::"n"(SP_EPI) : "memory"); }

// pooled rows of a unit that are complete once tiles 0..i are in the ring (conv row h is complete when its last
// valid column h*pitch + cols - 1 lies below 128*(i+1); pooled row ho needs conv rows 2ho-1 .. min(2ho+1, Ho-1))
[[maybe_unused]] __device__ __forceinline__ int sp_ready(const SpGeom& g, int cols, int i) {
    if (i < 0) return 0;
    const int P = 128 * (i + 1);
    if (P < cols) return 0;
    const int done = (P - cols) / g.pitch + 1;
    return done >= g.Ho ? g.Hp : (done >> 1);
}

// ---------------------------------------------------------------------------------------------------------------------
// shared by forward and backward: barrier layout, producer of the input patches, MMA issue loop
// ---------------------------------------------------------------------------------------------------------------------
struct SpBars {
    uint32_t base;
    __device__ uint32_t p_full(int b) const { return base + 8u * b; }
    __device__ uint32_t tm_full(int b) const { return base + 8u * (SP_NPB_MAX + b); }
    __device__ uint32_t tm_empty(int b) const { return base + 8u * (2 * SP_NPB_MAX + b); }
    __device__ uint32_t w_full() const { return base + 8u * (2 * SP_NPB_MAX + 2); }
    __device__ uint32_t d_full(int b) const { return base + 8u * (2 * SP_NPB_MAX + 3 + b); }
    __device__ uint32_t d_empty(int b) const { return base + 8u * (2 * SP_NPB_MAX + 5 + b); }
    __device__ uint32_t tmem_ptr() const { return base + 8u * (2 * SP_NPB_MAX + 7); }
    // fused-wgrad backward only
    __device__ uint32_t p_empty(int b) const { return base + 8u * (2 * SP_NPB_MAX + 8 + b); }
    __device__ uint32_t dy_full() const { return base + 8u * (3 * SP_NPB_MAX + 8); }
    __device__ uint32_t dy_empty() const { return base + 8u * (3 * SP_NPB_MAX + 9); }
    __device__ uint32_t acc_full() const { return base + 8u * (3 * SP_NPB_MAX + 10); }
    __device__ uint32_t acc_empty() const { return base + 8u * (3 * SP_NPB_MAX + 11); }
    static constexpr uint32_t BYTES = 8u * (3 * SP_NPB_MAX + 12);
};

__device__ __forceinline__ void sp_init_bars(const SpBars& B, const SpMaps& maps, int npb, bool backward) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.x_hi) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.x_lo) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.w) : "memory");
    if (backward) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.dout) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.didx) : "memory");
    }
    for (int b = 0; b < npb; ++b) { mbar_init(B.p_full(b), 1); mbar_init(B.tm_full(b), 1); }
    for (int b = 0; b < 2; ++b) {
        mbar_init(B.tm_empty(b), 4 * (SP_BN / SP_EC));
        mbar_init(B.d_full(b), 1);
        mbar_init(B.d_empty(b), 4 * (SP_BN / SP_EC));
    }
    mbar_init(B.w_full(), 1);
    if (backward) {
        for (int b = 0; b < npb; ++b) mbar_init(B.p_empty(b), 1);
        mbar_init(B.dy_full(), 4 * (SP_BN / SP_EC));
        mbar_init(B.dy_empty(), 1);
        mbar_init(B.acc_full(), 1);
        mbar_init(B.acc_empty(), 4 * (SP_BN / SP_EC));
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

// one elected thread: 16 taps x {X_hi x [W_hi ; W_lo] (N = 128), X_lo x W_hi (N = 64)} per tile, one commit per tile
__device__ __forceinline__ void sp_mma_loop(const SpGeom& g, const SpBars& B, uint32_t wbase, uint32_t pbase, uint32_t pbuf,
                                            uint32_t tmem_base, int my_tiles) {
    const uint32_t idesc_n = (1u << 4) | (1u << 7) | (1u << 10) | ((128u >> 4) << 24);
    const uint32_t idesc2 = idesc_n | ((uint32_t)((2 * SP_BN) >> 3) << 17);
    const uint32_t idesc1 = idesc_n | ((uint32_t)(SP_BN >> 3) << 17);
    uint32_t sh2[SP_TAPS];
#pragma unroll
    for (int t = 0; t < SP_TAPS; ++t) sh2[t] = (uint32_t)g.shift[t] * 2u;          // rows of 32 B in 16-byte units
    const uint32_t patch16 = (uint32_t)g.patch_bytes >> 4;
    const uint32_t w_lo32 = (wbase >> 4) | 0x10000u;
    const int NPB = g.npb;
    mbar_wait(B.w_full(), 0);
    for (int it = 0; it < my_tiles; ++it) {
        const int buf = it & 1, pb = it % NPB;
        const uint32_t td = tmem_base + (uint32_t)(buf * 2 * SP_BN), tcx = td + (uint32_t)SP_BN;
        const int i = it % g.tiles_per_unit;
        const int f0 = i * 128, hrow0 = f0 / g.pitch;
        const uint32_t a_lo32 = ((pbase + pb * pbuf + (uint32_t)(f0 - hrow0 * g.pitch) * 32u) >> 4) | 0x10000u;
        mbar_wait(B.tm_empty(buf), (((uint32_t)it >> 1) & 1u) ^ 1u);
        mbar_wait(B.p_full(pb), (uint32_t)(it / NPB) & 1u);
        tc_fence_after();
#pragma unroll
        for (int tap = 0; tap < SP_TAPS; ++tap) {
            const uint32_t ahi = a_lo32 + sh2[tap], alo = ahi + patch16, b = w_lo32 + (uint32_t)tap * (4096u >> 4);
            umma_bf16(td, SP_DHI | (uint64_t)ahi, SP_DHI | (uint64_t)b, idesc2, tap ? 1u : 0u);
            umma_bf16(tcx, SP_DHI | (uint64_t)alo, SP_DHI | (uint64_t)b, idesc1, 1u);
        }
        umma_commit(B.tm_full(pb));
    }
}

// =====================================================================================================================
// forward
// =====================================================================================================================
__global__ void __launch_bounds__(SP_THREADS, 1)
stem_pool_fwd_kernel(const __grid_constant__ SpMaps maps, const SpGeom g, const float* __restrict__ gamma,
                     float* __restrict__ ypool, uint8_t* __restrict__ idx, double* __restrict__ stats) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // smem: [filter bank 64 KB] [ring: 3 tiles x 128 rows x 272 B] [npb x (patch_hi | patch_lo)] [barriers] [BN partials] [sign]
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t ring = base + SP_W_BYTES;
    const uint32_t pbase = ring + (uint32_t)(RING_SLOTS * RING_SLOT);
    const uint32_t pbuf = 2u * (uint32_t)g.patch_bytes;
    const int NPB = g.npb;
    SpBars B;
    B.base = pbase + (uint32_t)NPB * pbuf;
    uint8_t* ring_g = smem_raw + (ring - smem_u32(smem_raw));
    float* stat_smem = reinterpret_cast<float*>(smem_raw + (B.base + SpBars::BYTES - smem_u32(smem_raw)));
    float* sgn = stat_smem + 128;
    if (threadIdx.x < 128) stat_smem[threadIdx.x] = 0.f;
    if (threadIdx.x < 64) sgn[threadIdx.x] = gamma[threadIdx.x] < 0.f ? -1.f : 1.f;
    // min-pool channels (gamma < 0) are rare: without any, the sign multiplications are skipped (CTA-uniform)
    const bool any_neg = __syncthreads_or(threadIdx.x < 64 && gamma[threadIdx.x] < 0.f) != 0;
    if (warp == 0 && lane == 0) sp_init_bars(B, maps, NPB, false);
    if (warp == 1) tmem_alloc(B.tmem_ptr(), 4u * SP_BN);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<const uint32_t*>(smem_raw + (B.tmem_ptr() - smem_u32(smem_raw)));
    const int my_units = ((int)blockIdx.x < g.total_units) ? (g.total_units - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
    const int my_tiles = my_units * g.tiles_per_unit;

    if (warp == 0) {
        if (elect_one() && my_tiles > 0) {
            mbar_expect_tx(B.w_full(), SP_W_BYTES);
            for (int j = 0; j < 8; ++j) tma_load_2d(&maps.w, base + j * 8192, B.w_full(), 0, j * 256);
            const uint32_t patch_tx = 2u * (uint32_t)(g.bhr * g.pitch) * 32u;
            for (int it = 0; it < my_tiles; ++it) {
                const int b = it % NPB;
                if (it >= NPB) mbar_wait(B.tm_full(b), ((uint32_t)(it / NPB) - 1u) & 1u);      // tile it-NPB retired: buffer free
                const TileCoord tc = sp_tile(g, it);
                mbar_expect_tx(B.p_full(b), patch_tx);
                tma_load_5d(&maps.x_hi, pbase + b * pbuf, B.p_full(b), 0, g.c0[tc.band] - 2, tc.hrow0 - 2, tc.t, tc.n);
                tma_load_5d(&maps.x_lo, pbase + b * pbuf + g.patch_bytes, B.p_full(b), 0, g.c0[tc.band] - 2, tc.hrow0 - 2, tc.t, tc.n);
            }
        }
    } else if (warp == 1) {
        if (elect_one() && my_tiles > 0) sp_mma_loop(g, B, base, pbase, pbuf, tmem_base, my_tiles);
    } else {
        // 16 epilogue warps = 4 TMEM lane quarters (warp % 4) x four 16-channel groups.  Per tile: accumulators ->
        // registers (+ BatchNorm partial sums) -> ring, pre-multiplied by sign(gamma); then the pooled rows the tile
        // completed are reduced from the ring by all 512 threads (one (pooled column, channel quad) each).
        // The epilogue is instruction-bound (ncu: 620 instructions per warp and tile in the first version, issue slots
        // 57 % busy, tensor pipe 37 %), so every per-tile index is incremental: no division or modulo in the tile loop.
        const int q = warp & 3, cg = (warp - 2) >> 2, c0 = cg * SP_EC;
        const int te = (int)threadIdx.x - 64, cq = te & 15, wo_t = te >> 4;
        const float4 sg = *reinterpret_cast<const float4*>(sgn + cq * 4);
        const int r = q * 32 + lane;
        const int pitch = g.pitch, Ho = g.Ho, Hp = g.Hp, Wp = g.Wp, tpu = g.tiles_per_unit;
        const int step_h = 128 / pitch, step_w = 128 - step_h * pitch;
        constexpr int RING_ROWS = RING_SLOTS * 128;
        const uint8_t* rb = ring_g + cq * 16;
        float rs[SP_EC], rq[SP_EC];
#pragma unroll
        for (int j = 0; j < SP_EC; ++j) { rs[j] = 0.f; rq[j] = 0.f; }
        int it = 0;
        for (int j = 0; j < my_units; ++j) {
            const int unit = (int)blockIdx.x + j * (int)gridDim.x;
            const int frame = unit / g.nbands, band = unit - frame * g.nbands;
            const int cols = g.cols[band], bc0 = g.c0[band], own0 = g.own0[band], npc = g.npc[band], wp0 = g.wp0[band];
            int h = r / pitch, wl = r - h * pitch;                     // this lane's conv position in tile 0 of the unit
            int ring_row = r;                                          // (tile % 3) * 128 + r
            int dq = (128 - cols) / pitch, drem = (128 - cols) - dq * pitch;     // complete conv rows - 1 after tile 0
            int ho = 0;                                                // next pooled row to reduce
            int mc = 0;                                                // ring row of conv row 2*ho, column 0 (mod RING_ROWS)
            float* yrow = ypool + ((size_t)frame * Hp * Wp << 6) + cq * 4;
            uint8_t* irow = idx + ((size_t)frame * Hp * Wp << 6) + cq * 4;
            for (int i = 0; i < tpu; ++i, ++it) {
                // ---- accumulators of tile `it` -> registers (the asynchronous TMEM loads and their wait stay adjacent:
                // no compiler-generated register copy may fall between them) ----
                uint32_t v[SP_EC], u[SP_EC];
                {
                    mbar_wait(B.tm_full(it % NPB), (uint32_t)(it / NPB) & 1u);
                    tc_fence_after();
                    const uint32_t td = tmem_base + (uint32_t)((it & 1) * 2 * SP_BN);
                    tmem_ld16_nowait(td + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
                    tmem_ld16_nowait(td + (uint32_t)SP_BN + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, u);
                }
                tmem_ld_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) sp_arrive(B.tm_empty(it & 1));
                const bool own = h < Ho && wl < cols && bc0 + wl >= own0;
                float4 o[SP_EC / 4];
#pragma unroll
                for (int jj = 0; jj < SP_EC / 4; ++jj) {
                    const float y0 = __uint_as_float(v[4 * jj]) + __uint_as_float(u[4 * jj]);
                    const float y1 = __uint_as_float(v[4 * jj + 1]) + __uint_as_float(u[4 * jj + 1]);
                    const float y2 = __uint_as_float(v[4 * jj + 2]) + __uint_as_float(u[4 * jj + 2]);
                    const float y3 = __uint_as_float(v[4 * jj + 3]) + __uint_as_float(u[4 * jj + 3]);
                    if (stats && own) {
                        rs[4 * jj] += y0; rs[4 * jj + 1] += y1; rs[4 * jj + 2] += y2; rs[4 * jj + 3] += y3;
                        rq[4 * jj] = fmaf(y0, y0, rq[4 * jj]); rq[4 * jj + 1] = fmaf(y1, y1, rq[4 * jj + 1]);
                        rq[4 * jj + 2] = fmaf(y2, y2, rq[4 * jj + 2]); rq[4 * jj + 3] = fmaf(y3, y3, rq[4 * jj + 3]);
                    }
                    o[jj] = make_float4(y0, y1, y2, y3);
                    if (any_neg) {
                        const float4 s4 = *reinterpret_cast<const float4*>(sgn + c0 + 4 * jj);
                        o[jj] = make_float4(y0 * s4.x, y1 * s4.y, y2 * s4.z, y3 * s4.w);
                    }
                }
                sp_epi_sync();                           // every thread is done reading the ring for the previous tile
                {
                    float4* dst = reinterpret_cast<float4*>(ring_g + ring_row * RING_PITCH + c0 * 4);
#pragma unroll
                    for (int jj = 0; jj < SP_EC / 4; ++jj) dst[jj] = o[jj];
                }
                sp_epi_sync();                           // the tile is in the ring
                // ---- pooled rows completed by this tile: conv rows 0..dq are complete ----
                const int ready = (dq + 1 >= Ho) ? Hp : ((dq + 1) >> 1);
                for (; ho < ready; ++ho) {
                    const bool up = ho > 0, dn = 2 * ho + 1 < Ho;
                    for (int wo = wo_t; wo < npc; wo += SP_EPI / 16) {
                        const int wg = wp0 + wo;
                        const bool lok = wg > 0 || bc0 > 0, rok = 2 * wg + 1 < g.Wo;       // window columns inside the frame
                        int m1 = mc + 2 * wg - bc0;                            // centre column of conv row 2*ho
                        if (m1 >= RING_ROWS) m1 -= RING_ROWS;
                        int m0 = m1 - pitch, m2 = m1 + pitch;
                        if (m0 < 0) m0 += RING_ROWS;
                        if (m2 >= RING_ROWS) m2 -= RING_ROWS;
                        float4 best = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
                        int kx = 4, ky = 4, kz = 4, kw = 4;
#define SP_CONSIDER(MM, KK)                                                                                   \
    {                                                                                                         \
        const float4 x = *reinterpret_cast<const float4*>(rb + (MM) * RING_PITCH);                            \
        kx = x.x > best.x ? (KK) : kx; ky = x.y > best.y ? (KK) : ky;                                         \
        kz = x.z > best.z ? (KK) : kz; kw = x.w > best.w ? (KK) : kw;                                         \
        best.x = fmaxf(best.x, x.x); best.y = fmaxf(best.y, x.y); best.z = fmaxf(best.z, x.z); best.w = fmaxf(best.w, x.w); \
    }
#define SP_ROW(MM, KB)                                                                                        \
    {                                                                                                         \
        const int ml = (MM) == 0 ? RING_ROWS - 1 : (MM) - 1, mr = (MM) == RING_ROWS - 1 ? 0 : (MM) + 1;       \
        if (lok) SP_CONSIDER(ml, KB)                                                                          \
        SP_CONSIDER(MM, (KB) + 1)                                                                             \
        if (rok) SP_CONSIDER(mr, (KB) + 2)                                                                    \
    }
#define SP_ROW_ALL(MM, KB)                                                                                    \
    {                                                                                                         \
        const int ml = (MM) == 0 ? RING_ROWS - 1 : (MM) - 1, mr = (MM) == RING_ROWS - 1 ? 0 : (MM) + 1;       \
        SP_CONSIDER(ml, KB)                                                                                   \
        SP_CONSIDER(MM, (KB) + 1)                                                                             \
        SP_CONSIDER(mr, (KB) + 2)                                                                             \
    }
                        if (up && dn && lok && rok) {                          // interior window: straight-line code
                            SP_ROW_ALL(m0, 0)
                            SP_ROW_ALL(m1, 3)
                            SP_ROW_ALL(m2, 6)
                        } else {
                            if (up) SP_ROW(m0, 0)
                            SP_ROW(m1, 3)
                            if (dn) SP_ROW(m2, 6)
                        }
#undef SP_ROW_ALL
#undef SP_ROW
#undef SP_CONSIDER
                        const size_t e = ((size_t)(ho * Wp + wg)) << 6;
                        if (any_neg) best = make_float4(best.x * sg.x, best.y * sg.y, best.z * sg.z, best.w * sg.w);
                        *reinterpret_cast<float4*>(yrow + e) = best;
                        *reinterpret_cast<uchar4*>(irow + e) = make_uchar4((unsigned char)kx, (unsigned char)ky, (unsigned char)kz, (unsigned char)kw);
                    }
                    mc += 2 * pitch;
                    if (mc >= RING_ROWS) mc -= RING_ROWS;
                }
                // next tile: this lane's position advances by 128 in padded-pitch order
                h += step_h; wl += step_w;
                if (wl >= pitch) { wl -= pitch; ++h; }
                dq += step_h; drem += step_w;
                if (drem >= pitch) { drem -= pitch; ++dq; }
                ring_row += 128;
                if (ring_row >= RING_ROWS) ring_row -= RING_ROWS;
            }
        }
        if (stats) {
            // transposing butterfly over the warp's 32 rows (as in stem_s2d_fwd_kernel)
#pragma unroll
            for (int off = 16, nn = SP_EC / 2; nn >= 1; off >>= 1, nn >>= 1) {
                const bool up = (lane & off) != 0;
#pragma unroll
                for (int i = 0; i < nn; ++i) {
                    const float s_send = up ? rs[i] : rs[i + nn], s_keep = up ? rs[i + nn] : rs[i];
                    const float q_send = up ? rq[i] : rq[i + nn], q_keep = up ? rq[i + nn] : rq[i];
                    rs[i] = s_keep + __shfl_xor_sync(0xffffffffu, s_send, off);
                    rq[i] = q_keep + __shfl_xor_sync(0xffffffffu, q_send, off);
                }
            }
            rs[0] += __shfl_xor_sync(0xffffffffu, rs[0], 1);
            rq[0] += __shfl_xor_sync(0xffffffffu, rq[0], 1);
            if ((lane & 1) == 0) {
                atomicAdd(&stat_smem[c0 + (lane >> 1)], rs[0]);
                atomicAdd(&stat_smem[64 + c0 + (lane >> 1)], rq[0]);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, 4u * SP_BN);
    if (stats && threadIdx.x < 128) atomicAdd(stats + threadIdx.x, (double)stat_smem[threadIdx.x]);
}

// ---- pooled grid: normalise + ReLU into operand planes; mark ReLU-dead windows -----------------------------------------
__device__ __forceinline__ void sp_store_planes(void* hi, void* lo, size_t off, float4 v) {
    const __nv_bfloat162 h01 = __floats2bfloat162_rn(v.x, v.y), h23 = __floats2bfloat162_rn(v.z, v.w);
    const float2 f01 = __bfloat1622float2(h01), f23 = __bfloat1622float2(h23);
    const __nv_bfloat162 l01 = __floats2bfloat162_rn(v.x - f01.x, v.y - f01.y);
    const __nv_bfloat162 l23 = __floats2bfloat162_rn(v.z - f23.x, v.w - f23.y);
    *reinterpret_cast<uint2*>(reinterpret_cast<uint16_t*>(hi) + off) =
        make_uint2(*reinterpret_cast<const uint32_t*>(&h01), *reinterpret_cast<const uint32_t*>(&h23));
    *reinterpret_cast<uint2*>(reinterpret_cast<uint16_t*>(lo) + off) =
        make_uint2(*reinterpret_cast<const uint32_t*>(&l01), *reinterpret_cast<const uint32_t*>(&l23));
}

__global__ void __launch_bounds__(256)
stem_pool_finalize_kernel(const float4* __restrict__ ypool, uchar4* __restrict__ idx, const float* __restrict__ mean,
                          const float* __restrict__ rstd, const float* __restrict__ gamma, const float* __restrict__ beta,
                          void* __restrict__ a_hi, void* __restrict__ a_lo, float4* __restrict__ a_rows, long long n4) {
    const int cq = threadIdx.x & 15;                                  // 256 threads: the channel quad is loop-invariant
    const float4 m = *reinterpret_cast<const float4*>(mean + cq * 4), r = *reinterpret_cast<const float4*>(rstd + cq * 4);
    const float4 ga = *reinterpret_cast<const float4*>(gamma + cq * 4), be = *reinterpret_cast<const float4*>(beta + cq * 4);
    const float4 sc = make_float4(ga.x * r.x, ga.y * r.y, ga.z * r.z, ga.w * r.w);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        const float4 y = ypool[i];
        float4 a = make_float4(fmaf(y.x - m.x, sc.x, be.x), fmaf(y.y - m.y, sc.y, be.y), fmaf(y.z - m.z, sc.z, be.z),
                               fmaf(y.w - m.w, sc.w, be.w));
        uchar4 k = idx[i];
        const bool dx = !(a.x > 0.f), dy = !(a.y > 0.f), dz = !(a.z > 0.f), dw = !(a.w > 0.f);
        a.x = dx ? 0.f : a.x; a.y = dy ? 0.f : a.y; a.z = dz ? 0.f : a.z; a.w = dw ? 0.f : a.w;
        if (dx | dy | dz | dw) {
            k.x = (unsigned char)((k.x & 0x0f) | (dx ? 0x80 : 0)); k.y = (unsigned char)((k.y & 0x0f) | (dy ? 0x80 : 0));
            k.z = (unsigned char)((k.z & 0x0f) | (dz ? 0x80 : 0)); k.w = (unsigned char)((k.w & 0x0f) | (dw ? 0x80 : 0));
            idx[i] = k;
        }
        sp_store_planes(a_hi, a_lo, (size_t)i * 4, a);
        if (a_rows) a_rows[i] = a;
    }
}

// ---- bn1 backward sums on the pooled grid: sum g | sum g * xhat with g = dout * [alive], xhat from the kept conv value ----
__global__ void __launch_bounds__(256)
stem_pool_bwd_reduce_kernel(const float4* __restrict__ ypool, const float4* __restrict__ dout, const uchar4* __restrict__ idx,
                            const float* __restrict__ mean, const float* __restrict__ rstd, long long n4,
                            double* __restrict__ ws) {
    __shared__ double red[16][16 * 8];
    const int cq = threadIdx.x & 15, rl = threadIdx.x >> 4;
    const float4 m = *reinterpret_cast<const float4*>(mean + cq * 4), r = *reinterpret_cast<const float4*>(rstd + cq * 4);
    double s[4] = {0, 0, 0, 0}, sx[4] = {0, 0, 0, 0};
    const long long stride = (long long)gridDim.x * blockDim.x;
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    while (i < n4) {
        float4 pa = make_float4(0.f, 0.f, 0.f, 0.f), pb = pa;
        for (int it = 0; it < 32 && i < n4; ++it, i += stride) {        // fp32 over <= 32 rows, then fp64
            const float4 y = ypool[i], d = dout[i];
            const uchar4 k = idx[i];
            const float gx = (k.x & 0x80) ? 0.f : d.x, gy = (k.y & 0x80) ? 0.f : d.y;
            const float gz = (k.z & 0x80) ? 0.f : d.z, gw = (k.w & 0x80) ? 0.f : d.w;
            pa.x += gx; pa.y += gy; pa.z += gz; pa.w += gw;
            pb.x = fmaf(gx, (y.x - m.x) * r.x, pb.x); pb.y = fmaf(gy, (y.y - m.y) * r.y, pb.y);
            pb.z = fmaf(gz, (y.z - m.z) * r.z, pb.z); pb.w = fmaf(gw, (y.w - m.w) * r.w, pb.w);
        }
        s[0] += pa.x; s[1] += pa.y; s[2] += pa.z; s[3] += pa.w;
        sx[0] += pb.x; sx[1] += pb.y; sx[2] += pb.z; sx[3] += pb.w;
    }
#pragma unroll
    for (int e = 0; e < 4; ++e) { red[rl][cq * 8 + e] = s[e]; red[rl][cq * 8 + 4 + e] = sx[e]; }
    __syncthreads();
    if (threadIdx.x < 128) {
        double a = 0.0;
        for (int k = 0; k < 16; ++k) a += red[k][threadIdx.x];
        const int q = threadIdx.x >> 3, e = threadIdx.x & 7;
        atomicAdd(ws + (e < 4 ? 0 : 64) + q * 4 + (e & 3), a);
    }
}

__global__ void stem_pool_bwd_finalize_kernel(const double* __restrict__ ws, float* __restrict__ dgamma, float* __restrict__ dbeta) {
    const int c = threadIdx.x;
    if (c < 64) { dbeta[c] = (float)ws[c]; dgamma[c] = (float)ws[64 + c]; }
}

// =====================================================================================================================
// backward: recompute conv1, gather the pooled gradient, BatchNorm backward; the gradient tile either goes to HBM as
// split-bf16 planes (FUSE = false: the separate conv1 wgrad kernel reads them) or stays in shared memory as the MN-major
// operand of the conv1 wgrad MMAs issued by the same kernel (FUSE = true: the 5.4 GB gradient is never materialised).
// =====================================================================================================================
constexpr int SP_WG_CHAIN = 64;            // tiles per in-TMEM wgrad accumulation chain (1024 truncating accumulations)
constexpr uint32_t SP_DY_PLANE = 128 * 128; // one plane of the gradient tile: 128 positions x 64 bf16

// pooled gradient of one conv position (h, wc), 16 channels: sum over the (<= 4) pooling windows that contain it of
// dout[window] * [window selected this position]; `st` = TMA-staged window rows (dout channel halves | index).
// One pass per candidate window: 5 shared-memory loads + 3 instructions per channel (mask, compare, predicated add).
__device__ __forceinline__ void sp_gather_window(float (&gg)[SP_EC], const uint8_t* st_d, const uint8_t* st_i, int wrow, uint32_t kk,
                                                 int cg, int cgl) {
    const uint4 iv = *reinterpret_cast<const uint4*>(st_i + wrow * 64 + ((cg ^ ((wrow >> 1) & 3)) << 4));
    const uint32_t iw[4] = {iv.x, iv.y, iv.z, iv.w};
    const uint8_t* drow = st_d + wrow * 128;
    const int x7 = (wrow & 7) << 4;
    const uint32_t k0 = kk, k1 = kk << 8, k2 = kk << 16, k3 = kk << 24;
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) {
        const float4 d = *reinterpret_cast<const float4*>(drow + (((cgl * 4 + jj) << 4) ^ x7));
        const uint32_t w4 = iw[jj];
        if ((w4 & 0x000000ffu) == k0) gg[4 * jj] += d.x;
        if ((w4 & 0x0000ff00u) == k1) gg[4 * jj + 1] += d.y;
        if ((w4 & 0x00ff0000u) == k2) gg[4 * jj + 2] += d.z;
        if ((w4 & 0xff000000u) == k3) gg[4 * jj + 3] += d.w;
    }
}

__device__ __forceinline__ void sp_gather(float (&gg)[SP_EC], const uint8_t* st, int dhalf_bytes, int bc, int Hp, int Wp,
                                          int h, int wc, int ho_box0, int wq0, int cg) {
    const int cgl = cg & 1;
    const uint8_t* st_d = st + (cg >> 1) * dhalf_bytes;
    const uint8_t* st_i = st + 2 * dhalf_bytes;
    const bool hodd = (h & 1) != 0, wodd = (wc & 1) != 0;
    const int ho0 = h >> 1, wo0 = wc >> 1;
    const int w00 = (ho0 - ho_box0) * bc + (wo0 - wq0);
    // window (a, b) = pooled (ho0 + a, wo0 + b); the position's index inside it: dh = odd ? (a ? 0 : 2) : 1, dw likewise
    const uint32_t dh_a0 = hodd ? 2u : 1u, dw_b0 = wodd ? 2u : 1u;
    if (ho0 < Hp) {
        if (wo0 < Wp) sp_gather_window(gg, st_d, st_i, w00, dh_a0 * 3u + dw_b0, cg, cgl);
        if (wodd && wo0 + 1 < Wp) sp_gather_window(gg, st_d, st_i, w00 + 1, dh_a0 * 3u, cg, cgl);
    }
    if (hodd && ho0 + 1 < Hp) {
        if (wo0 < Wp) sp_gather_window(gg, st_d, st_i, w00 + bc, dw_b0, cg, cgl);
        if (wodd && wo0 + 1 < Wp) sp_gather_window(gg, st_d, st_i, w00 + bc + 1, 0u, cg, cgl);
    }
}

// dy = k*(g - mb - xhat*mg) = k*g - P*y - Q (P = k*mg*rstd, Q = k*mb - P*mean) for 16 channels -> packed hi / lo bf16
__device__ __forceinline__ void sp_dy_planes(uint32_t (&ph)[SP_EC / 2], uint32_t (&pl)[SP_EC / 2], const float (&yy)[SP_EC],
                                             const float (&gg)[SP_EC], const float4* kq, const float4* pq, const float4* qq) {
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) {
        const float4 kk = kq[jj], pp = pq[jj], qv = qq[jj];
        const float kv[4] = {kk.x, kk.y, kk.z, kk.w}, pv[4] = {pp.x, pp.y, pp.z, pp.w}, qw[4] = {qv.x, qv.y, qv.z, qv.w};
        float d4[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            d4[e] = fmaf(kv[e], gg[4 * jj + e], fmaf(-pv[e], yy[4 * jj + e], -qw[e]));
        }
#pragma unroll
        for (int e = 0; e < 4; e += 2) {
            const __nv_bfloat162 hh = __floats2bfloat162_rn(d4[e], d4[e + 1]);
            const float2 hf = __bfloat1622float2(hh);
            const __nv_bfloat162 ll = __floats2bfloat162_rn(d4[e] - hf.x, d4[e + 1] - hf.y);
            ph[2 * jj + (e >> 1)] = *reinterpret_cast<const uint32_t*>(&hh);
            pl[2 * jj + (e >> 1)] = *reinterpret_cast<const uint32_t*>(&ll);
        }
    }
}

// fused schedule of the MMA thread:  conv(0); for every tile: conv(it+1); wgrad(it) once the epilogue has written the
// gradient tile.  wgrad (as in stem_s2d_wgrad_kernel): K = positions, A (M = 128) = [dY_hi^T ; dY_lo^T] (MN-major
// SWIZZLE_128B, the two planes LBO apart), B (N = 64) = X2^T of the four taps of a filter row (MN-major SWIZZLE_32B over
// the conv's own input patch, one patch row apart); 4 accumulators of 64 columns, one per filter row.
__device__ __forceinline__ void sp_mma_loop_fused(const SpGeom& g, const SpBars& B, uint32_t wbase, uint32_t pbase, uint32_t pbuf,
                                                  uint32_t dybase, uint32_t tmem_base, int my_tiles) {
    const uint32_t idesc_n = (1u << 4) | (1u << 7) | (1u << 10) | ((128u >> 4) << 24);
    const uint32_t idesc2 = idesc_n | ((uint32_t)((2 * SP_BN) >> 3) << 17);
    const uint32_t idesc1 = idesc_n | ((uint32_t)(SP_BN >> 3) << 17);
    const uint32_t idesc_w = idesc_n | (1u << 15) | (1u << 16) | ((64u >> 3) << 17);      // both operands MN-major, N = 64
    constexpr uint64_t DHI128 = 0x40004040ull << 32;                                     // SWIZZLE_128B, SBO = 1024 B
    uint32_t sh2[SP_TAPS];
#pragma unroll
    for (int t = 0; t < SP_TAPS; ++t) sh2[t] = (uint32_t)g.shift[t] * 2u;
    const uint32_t patch16 = (uint32_t)g.patch_bytes >> 4;
    const uint32_t w_lo32 = (wbase >> 4) | 0x10000u;
    const uint32_t y16 = (dybase >> 4) | ((SP_DY_PLANE >> 4) << 16);                     // LBO = plane pitch
    const uint32_t arow16 = (uint32_t)g.pitch * 2u;
    const uint32_t tmem_w = tmem_base + 4u * SP_BN;
    const int NPB = g.npb, tpu = g.tiles_per_unit;
    mbar_wait(B.w_full(), 0);
    auto conv = [&](int it) {
        const int buf = it & 1, pb = it % NPB;
        const uint32_t td = tmem_base + (uint32_t)(buf * 2 * SP_BN), tcx = td + (uint32_t)SP_BN;
        const int i = it % tpu;
        const int f0 = i * 128, hrow0 = f0 / g.pitch;
        const uint32_t a_lo32 = ((pbase + pb * pbuf + (uint32_t)(f0 - hrow0 * g.pitch) * 32u) >> 4) | 0x10000u;
        mbar_wait(B.tm_empty(buf), (((uint32_t)it >> 1) & 1u) ^ 1u);
        mbar_wait(B.p_full(pb), (uint32_t)(it / NPB) & 1u);
        tc_fence_after();
#pragma unroll
        for (int tap = 0; tap < SP_TAPS; ++tap) {
            const uint32_t ahi = a_lo32 + sh2[tap], alo = ahi + patch16, b = w_lo32 + (uint32_t)tap * (4096u >> 4);
            umma_bf16(td, SP_DHI | (uint64_t)ahi, SP_DHI | (uint64_t)b, idesc2, tap ? 1u : 0u);
            umma_bf16(tcx, SP_DHI | (uint64_t)alo, SP_DHI | (uint64_t)b, idesc1, 1u);
        }
        umma_commit(B.tm_full(buf));
    };
    if (my_tiles > 0) conv(0);
    for (int it = 0; it < my_tiles; ++it) {
        if (it + 1 < my_tiles) conv(it + 1);
        const int in_chain = it % SP_WG_CHAIN, pb = it % NPB;
        const int i = it % tpu;
        const int f0 = i * 128, hrow0 = f0 / g.pitch;
        if (in_chain == 0 && it > 0) mbar_wait(B.acc_empty(), ((uint32_t)(it / SP_WG_CHAIN) - 1u) & 1u);
        mbar_wait(B.dy_full(), (uint32_t)it & 1u);
        tc_fence_after();
        const uint32_t x16 = (((pbase + pb * pbuf) >> 4) + (uint32_t)(f0 - hrow0 * g.pitch) * 2u) | (2u << 16);   // LBO = one patch row
#pragma unroll
        for (int k = 0; k < 8; ++k) {                      // UMMA_K = 16 positions: 2048 B of gradient rows, 512 B of X2 rows
            const uint64_t ya = DHI128 | (uint64_t)(y16 + k * 128);
#pragma unroll
            for (int a = 0; a < 4; ++a) {
                const uint32_t xa = x16 + (uint32_t)a * arow16 + k * 32;
                const uint32_t d = tmem_w + (uint32_t)(a * 64);
                umma_bf16(d, ya, SP_DHI | (uint64_t)xa, idesc_w, (in_chain | k) ? 1u : 0u);
                umma_bf16(d, ya, SP_DHI | (uint64_t)(xa + patch16), idesc_w, 1u);
            }
        }
        umma_commit(B.p_empty(pb));                        // the input patch is free
        umma_commit(B.dy_empty());                         // the gradient tile is free
        if (in_chain == SP_WG_CHAIN - 1 || it == my_tiles - 1) umma_commit(B.acc_full());
    }
}

template <bool FUSE>
__global__ void __launch_bounds__(SP_THREADS, 1)
stem_pool_bwd_kernel(const __grid_constant__ SpMaps maps, const SpGeom g, const float* __restrict__ mean,
                     const float* __restrict__ rstd, const float* __restrict__ gamma, const double* __restrict__ ws,
                     uint16_t* __restrict__ dy_hi, uint16_t* __restrict__ dy_lo, float* __restrict__ dw) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // smem: [filter bank 64 KB] [npb x (patch_hi | patch_lo)] [nds x (dout ch 0-31 | dout ch 32-63 | idx)]
    //       [FUSE: gradient tile hi | lo, 32 KB] [barriers] [constants]
    constexpr int NDS = FUSE ? 1 : 2;
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t pbase = base + SP_W_BYTES;
    const uint32_t pbuf = 2u * (uint32_t)g.patch_bytes;
    const int NPB = g.npb;
    const uint32_t dbase = pbase + (uint32_t)NPB * pbuf;
    const uint32_t dstage = 2u * (uint32_t)g.dhalf_bytes + (uint32_t)g.didx_bytes;
    const uint32_t dybase = dbase + NDS * dstage;
    SpBars B;
    B.base = dybase + (FUSE ? 2u * SP_DY_PLANE : 0u);
    const uint8_t* d_g = smem_raw + (dbase - smem_u32(smem_raw));
    uint8_t* dy_g = smem_raw + (dybase - smem_u32(smem_raw));
    float* cst = reinterpret_cast<float*>(smem_raw + (B.base + SpBars::BYTES - smem_u32(smem_raw)));   // k | P | Q
    if (threadIdx.x < 64) {
        const int c = threadIdx.x;
        const double k = (double)gamma[c] * (double)rstd[c];
        const double mb = ws[c] * g.inv_n, mg = ws[64 + c] * g.inv_n;
        const double Pc = k * mg * (double)rstd[c];
        cst[c] = (float)k;
        cst[64 + c] = (float)Pc;
        cst[128 + c] = (float)(k * mb - Pc * (double)mean[c]);
    }
    constexpr uint32_t TMEM_COLS = FUSE ? 512u : 4u * SP_BN;
    if (warp == 0 && lane == 0) sp_init_bars(B, maps, NPB, true);
    if (warp == 1) tmem_alloc(B.tmem_ptr(), TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<const uint32_t*>(smem_raw + (B.tmem_ptr() - smem_u32(smem_raw)));
    const int my_units = ((int)blockIdx.x < g.total_units) ? (g.total_units - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
    const int my_tiles = my_units * g.tiles_per_unit;

    if (warp == 0) {
        if (elect_one() && my_tiles > 0) {
            mbar_expect_tx(B.w_full(), SP_W_BYTES);
            for (int j = 0; j < 8; ++j) tma_load_2d(&maps.w, base + j * 8192, B.w_full(), 0, j * 256);
            const uint32_t patch_tx = 2u * (uint32_t)(g.bhr * g.pitch) * 32u;
            const uint32_t d_tx = (uint32_t)(g.bc * g.br) * (2u * 128u + 64u);
            for (int it = 0; it < my_tiles; ++it) {
                const int b = it % NPB, ds = FUSE ? 0 : (it & 1);
                const TileCoord tc = sp_tile(g, it);
                if (it >= NPB) {
                    if (FUSE) mbar_wait(B.p_empty(b), ((uint32_t)(it / NPB) - 1u) & 1u);       // wgrad of tile it-NPB retired
                    else mbar_wait(B.tm_full(b), ((uint32_t)(it / NPB) - 1u) & 1u);            // conv of tile it-NPB retired
                }
                mbar_expect_tx(B.p_full(b), patch_tx);
                tma_load_5d(&maps.x_hi, pbase + b * pbuf, B.p_full(b), 0, g.c0[tc.band] - 2, tc.hrow0 - 2, tc.t, tc.n);
                tma_load_5d(&maps.x_lo, pbase + b * pbuf + g.patch_bytes, B.p_full(b), 0, g.c0[tc.band] - 2, tc.hrow0 - 2, tc.t, tc.n);
                if (it >= NDS) mbar_wait(B.d_empty(ds), (uint32_t)(it / NDS - 1) & 1u);
                mbar_expect_tx(B.d_full(ds), d_tx);
                const uint32_t sd = dbase + ds * dstage;
                const int ho0 = tc.hrow0 >> 1, wq0 = g.wp0[tc.band];
                tma_load_5d(&maps.dout, sd, B.d_full(ds), 0, wq0, ho0, tc.t, tc.n);
                tma_load_5d(&maps.dout, sd + g.dhalf_bytes, B.d_full(ds), 32, wq0, ho0, tc.t, tc.n);
                tma_load_5d(&maps.didx, sd + 2 * g.dhalf_bytes, B.d_full(ds), 0, wq0, ho0, tc.t, tc.n);
            }
        }
    } else if (warp == 1) {
        if (elect_one() && my_tiles > 0) {
            if (FUSE) sp_mma_loop_fused(g, B, base, pbase, pbuf, dybase, tmem_base, my_tiles);
            else sp_mma_loop(g, B, base, pbase, pbuf, tmem_base, my_tiles);
        }
    } else {
        // 16 epilogue warps: lane = conv position of the tile, warp = (TMEM lane quarter, 16-channel group).  Instruction-
        // bound (ncu), hence: incremental position arithmetic (no divisions per tile), byte-parallel index compares
        // (vcmpeq4 + sign-replicating prmt), the BatchNorm backward folded into two FMAs per element.
        const int q = warp & 3, cg = (warp - 2) >> 2, c0 = cg * SP_EC;
        const int r = q * 32 + lane;
        const int pitch = g.pitch, Ho = g.Ho, Wo = g.Wo, Hp = g.Hp, Wp = g.Wp, bc = g.bc, dhalf = g.dhalf_bytes, tpu = g.tiles_per_unit;
        const int step_h = 128 / pitch, step_w = 128 - step_h * pitch;
        const float4* kq = reinterpret_cast<const float4*>(cst + c0);
        const float4* pq = reinterpret_cast<const float4*>(cst + 64 + c0);
        const float4* qq = reinterpret_cast<const float4*>(cst + 128 + c0);
        int it = 0;
        for (int j = 0; j < my_units; ++j) {
            const int unit = (int)blockIdx.x + j * (int)gridDim.x;
            const int frame = unit / g.nbands, band = unit - frame * g.nbands;
            const int cols = g.cols[band], bc0 = g.c0[band], own0 = g.own0[band], wq0 = g.wp0[band];
            int h = r / pitch, wl = r - h * pitch;                     // this lane's conv position in tile 0 of the unit
            int hb = 0, wb = 0;                                        // position of the tile's row 0 (-> first conv row)
            for (int i = 0; i < tpu; ++i, ++it) {
                const int ds = FUSE ? 0 : (it & 1);
                const int wc = bc0 + wl;
                const bool own = h < Ho && wl < cols && wc >= own0;
                float yy[SP_EC];
                {
                    uint32_t v[SP_EC], u[SP_EC];
                    if (FUSE) mbar_wait(B.tm_full(it & 1), ((uint32_t)it >> 1) & 1u);
                    else mbar_wait(B.tm_full(it % NPB), (uint32_t)(it / NPB) & 1u);
                    tc_fence_after();
                    const uint32_t td = tmem_base + (uint32_t)((it & 1) * 2 * SP_BN);
                    tmem_ld16_nowait(td + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
                    tmem_ld16_nowait(td + (uint32_t)SP_BN + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, u);
                    tmem_ld_wait();
#pragma unroll
                    for (int jj = 0; jj < SP_EC; ++jj) yy[jj] = __uint_as_float(v[jj]) + __uint_as_float(u[jj]);
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) sp_arrive(B.tm_empty(it & 1));
                mbar_wait(B.d_full(ds), (uint32_t)(it / NDS) & 1u);
                float gg[SP_EC];
#pragma unroll
                for (int jj = 0; jj < SP_EC; ++jj) gg[jj] = 0.f;
                if (own) sp_gather(gg, d_g + ds * dstage, dhalf, bc, Hp, Wp, h, wc, hb >> 1, wq0, cg);
                __syncwarp();
                if (lane == 0) sp_arrive(B.d_empty(ds));
                uint32_t ph[SP_EC / 2], pl[SP_EC / 2];
                if (own) {
                    sp_dy_planes(ph, pl, yy, gg, kq, pq, qq);
                } else {
#pragma unroll
                    for (int jj = 0; jj < SP_EC / 2; ++jj) { ph[jj] = 0u; pl[jj] = 0u; }
                }
                if (FUSE) {
                    // gradient tile in shared memory: row = position, 64 channels x bf16 = 128 B, 128B-swizzled (the layout a
                    // SWIZZLE_128B TMA box would have); rows that are not owned conv outputs are zero
                    if (it > 0) mbar_wait(B.dy_empty(), ((uint32_t)it - 1u) & 1u);             // wgrad(it-1) has read the tile
                    uint8_t* rowp = dy_g + r * 128;
                    const int x7 = r & 7;
                    *reinterpret_cast<uint4*>(rowp + (((2 * cg) ^ x7) << 4)) = make_uint4(ph[0], ph[1], ph[2], ph[3]);
                    *reinterpret_cast<uint4*>(rowp + (((2 * cg + 1) ^ x7) << 4)) = make_uint4(ph[4], ph[5], ph[6], ph[7]);
                    *reinterpret_cast<uint4*>(rowp + SP_DY_PLANE + (((2 * cg) ^ x7) << 4)) = make_uint4(pl[0], pl[1], pl[2], pl[3]);
                    *reinterpret_cast<uint4*>(rowp + SP_DY_PLANE + (((2 * cg + 1) ^ x7) << 4)) = make_uint4(pl[4], pl[5], pl[6], pl[7]);
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");              // generic-proxy writes -> UMMA reads
                    __syncwarp();
                    if (lane == 0) sp_arrive(B.dy_full());
                    if ((it % SP_WG_CHAIN) == SP_WG_CHAIN - 1 || it == my_tiles - 1) {
                        // drain the wgrad accumulators: TMEM lane = (plane of dY, co), column = (b, ch) of filter row a
                        mbar_wait(B.acc_full(), (uint32_t)(it / SP_WG_CHAIN) & 1u);
                        tc_fence_after();
                        const int co = r & 63;
                        float* dwc = dw + (size_t)co * 147;
#pragma unroll
                        for (int a = 0; a < 4; ++a) {
                            uint32_t wv[SP_EC];
                            tmem_ld16_nowait(tmem_base + 4u * SP_BN + ((uint32_t)(q * 32) << 16) + (uint32_t)(a * 64 + c0), wv);
                            tmem_ld_wait();
#pragma unroll
                            for (int ch = 0; ch < 12; ++ch) {
                                const int cc = ch >> 2, rr = (ch >> 1) & 1, ss = ch & 1;
                                const int kh = 2 * a + rr - 1, kw = 2 * cg + ss - 1;
                                if (kh >= 0 && kw >= 0) atomicAdd(dwc + cc * 49 + kh * 7 + kw, __uint_as_float(wv[ch]));
                            }
                        }
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) sp_arrive(B.acc_empty());
                    }
                } else if (own) {
                    const size_t e = (((((size_t)frame * Ho + h) * Wo + wc)) << 6) + c0;
                    uint4* dh4 = reinterpret_cast<uint4*>(dy_hi + e);
                    uint4* dl4 = reinterpret_cast<uint4*>(dy_lo + e);
                    dh4[0] = make_uint4(ph[0], ph[1], ph[2], ph[3]); dh4[1] = make_uint4(ph[4], ph[5], ph[6], ph[7]);
                    dl4[0] = make_uint4(pl[0], pl[1], pl[2], pl[3]); dl4[1] = make_uint4(pl[4], pl[5], pl[6], pl[7]);
                }
                h += step_h; wl += step_w;
                if (wl >= pitch) { wl -= pitch; ++h; }
                hb += step_h; wb += step_w;
                if (wb >= pitch) { wb -= pitch; ++hb; }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, TMEM_COLS);
}

// ---------------------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------------------
PFN_cuTensorMapEncodeTiled_v12000 sp_encode() {
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

size_t sp_fwd_smem(const SpGeom& g) {
    return 1024 + SP_W_BYTES + (size_t)RING_SLOTS * RING_SLOT + (size_t)g.npb * 2 * g.patch_bytes + SpBars::BYTES + 4 * (128 + 64);
}
size_t sp_bwd_smem(const SpGeom& g, bool fuse = false) {
    return 1024 + SP_W_BYTES + (size_t)g.npb * 2 * g.patch_bytes + (fuse ? 1 : 2) * (2 * (size_t)g.dhalf_bytes + g.didx_bytes) +
           (fuse ? 2 * SP_DY_PLANE : 0) + SpBars::BYTES + 4 * 192;
}

// frame geometry: column bands, pitch, patch box, tiles; returns an error string or nullptr
const char* sp_geometry(SpGeom& g, int NB, int T, int H, int W) {
    memset(&g, 0, sizeof(g));
    if ((H & 1) || (W & 1)) return "H and W must be even";
    g.Ho = H / 2; g.Wo = W / 2; g.T = T;
    g.Hp = (g.Ho - 1) / 2 + 1; g.Wp = (g.Wo - 1) / 2 + 1;
    const size_t budget = 227 * 1024;
    for (int nb = 1; nb <= SP_MAX_BANDS; ++nb) {
        if (nb > g.Wp) break;
        const int per = (g.Wp + nb - 1) / nb;
        int maxcols = 0, maxnpc = 0, prev_c1 = 0;
        bool ok = true;
        for (int b = 0; b < nb; ++b) {
            const int wp0 = b * per, wp1 = (wp0 + per < g.Wp) ? wp0 + per : g.Wp;
            if (wp1 <= wp0) { ok = false; break; }
            const int c0 = (2 * wp0 - 1 > 0) ? 2 * wp0 - 1 : 0;
            const int c1 = (2 * (wp1 - 1) + 2 < g.Wo) ? 2 * (wp1 - 1) + 2 : g.Wo;
            g.wp0[b] = wp0; g.npc[b] = wp1 - wp0; g.c0[b] = c0; g.cols[b] = c1 - c0; g.own0[b] = b ? prev_c1 : 0;
            prev_c1 = c1;
            if (c1 - c0 > maxcols) maxcols = c1 - c0;
            if (wp1 - wp0 > maxnpc) maxnpc = wp1 - wp0;
        }
        if (!ok) continue;
        g.nbands = nb;
        g.pitch = maxcols + 3;
        // a pooling window (3 conv rows of one band) must span <= 3 tiles of the ring
        if (2 * g.pitch + maxcols > 257) continue;
        g.bhr = 4 + (130 + g.pitch - 1) / g.pitch;                 // rows [rowoff, rowoff + 128 + 3*pitch + 3), rowoff < pitch
        if (g.pitch > 256 || g.bhr > 256) continue;
        g.patch_bytes = ((g.bhr * g.pitch * 32 + 1023) / 1024) * 1024;
        // backward window box: pooled rows touched by one tile, pooled columns of a band + 1
        const int R = (g.pitch - 1 + 127) / g.pitch + 1;
        g.br = (R + 1) / 2 + 1;
        g.bc = maxnpc + 1;
        if (g.bc > 256 || g.br > 256) continue;
        g.dhalf_bytes = ((g.br * g.bc * 128 + 1023) / 1024) * 1024;
        g.didx_bytes = ((g.br * g.bc * 64 + 1023) / 1024) * 1024;
        g.npb = 2;
        if (sp_fwd_smem(g) > budget || sp_bwd_smem(g) > budget) continue;
        g.tiles_per_unit = (g.Ho * g.pitch + 127) / 128;
        const long long units = (long long)NB * T * nb;
        if (units * g.tiles_per_unit >= (1ll << 31)) return "too many tiles";
        g.total_units = (int)units;
        for (int a = 0; a < 4; ++a)
            for (int b = 0; b < 4; ++b) g.shift[a * 4 + b] = a * g.pitch + b;
        g.inv_n = 1.0 / ((double)NB * T * g.Ho * g.Wo);
        return nullptr;
    }
    return "frame does not fit the pooled-stem schedule";
}

int sp_x_maps(SpMaps& maps, const SpGeom& g, const void* x2_hi, const void* x2_lo, const void* wp, int NB) {
    auto enc = sp_encode();
    DPC_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled entry point not available");
    const cuuint32_t es[5] = {1, 1, 1, 1, 1};
    const cuuint64_t gd[5] = {SP_CH, (cuuint64_t)g.Wo, (cuuint64_t)g.Ho, (cuuint64_t)g.T, (cuuint64_t)NB};
    const cuuint64_t gs[4] = {32, (cuuint64_t)g.Wo * 32, (cuuint64_t)g.Ho * g.Wo * 32, (cuuint64_t)g.T * g.Ho * g.Wo * 32};
    const cuuint32_t bx[5] = {SP_CH, (cuuint32_t)g.pitch, (cuuint32_t)g.bhr, 1, 1};
    for (int i = 0; i < 2; ++i) {
        CUresult r = enc(i ? &maps.x_lo : &maps.x_hi, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(i ? x2_lo : x2_hi),
                         gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_32B,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        DPC_REQUIRE(r == CUDA_SUCCESS, "stem_pool: cuTensorMapEncodeTiled (x) failed (%d)", (int)r);
    }
    const cuuint64_t wd[2] = {SP_CH, (cuuint64_t)SP_TAPS * 2 * SP_BN};
    const cuuint64_t wst[1] = {32};
    const cuuint32_t wb[2] = {SP_CH, 256}, we[2] = {1, 1};
    CUresult r = enc(&maps.w, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(wp), wd, wst, wb, we,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    DPC_REQUIRE(r == CUDA_SUCCESS, "stem_pool: cuTensorMapEncodeTiled (w) failed (%d)", (int)r);
    return DPC_OK;
}

}  // namespace

// 0: frames of H x W are not handled by the pooled-stem kernels (odd extents, bands that do not fit shared memory);
// 1: handled; 2: handled, and the backward can also run conv1's wgrad in the same kernel (dpc_stem_pool_bwd_wgrad)
extern "C" int dpc_stem_pool_supported(int H, int W) {
    SpGeom g;
    if (H <= 0 || W <= 0 || sp_geometry(g, 1, 1, H, W) != nullptr) return 0;
    g.npb = 3;
    return sp_bwd_smem(g, true) <= 227 * 1024 ? 2 : 1;
}

// x2 planes [NB,T,H/2,W/2,16] (dpc_stem_s2d_pack) + packed filter bank `wp` (dpc_stem_s2d_wpack) ->
//   ypool [NB,T,Hp,Wp,64] fp32: per pooled position the conv1 value that max-pool o relu o bn1 selects,
//   idx   [NB,T,Hp,Wp,64] uint8: its index dh*3 + dw inside the 3x3 window,
//   bn_ws 128 doubles: per-channel sum | sum of squares of conv1 over ALL H/2 x W/2 positions.
extern "C" int dpc_stem_pool_fwd(const void* x2_hi, const void* x2_lo, const void* wp, const float* gamma, float* ypool,
                                 void* idx, double* bn_ws, int NB, int T, int H, int W, void* stream) {
    DPC_REQUIRE(x2_hi && x2_lo && wp && gamma && ypool && idx && NB > 0 && T > 0 && H > 0 && W > 0, "dpc_stem_pool_fwd: bad args");
    SpGeom g;
    if (const char* e = sp_geometry(g, NB, T, H, W)) { dpc_set_error("dpc_stem_pool_fwd: %s (H %d, W %d)", e, H, W); return DPC_ERR_UNSUPPORTED; }
    cudaStream_t st = as_stream(stream);
    while (g.npb < SP_NPB_MAX) { ++g.npb; if (sp_fwd_smem(g) > 227 * 1024) { --g.npb; break; } }
    SpMaps maps;
    memset(&maps, 0, sizeof(maps));
    if (int rc = sp_x_maps(maps, g, x2_hi, x2_lo, wp, NB)) return rc;
    const size_t smem = sp_fwd_smem(g);
    if (bn_ws) DPC_CUDA(cudaMemsetAsync(bn_ws, 0, sizeof(double) * 128, st));
    DPC_CUDA(cudaFuncSetAttribute(stem_pool_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int sms = dpc_num_sms();
    const int grid = g.total_units < sms ? g.total_units : sms;
    stem_pool_fwd_kernel<<<grid, SP_THREADS, smem, st>>>(maps, g, gamma, ypool, reinterpret_cast<uint8_t*>(idx), bn_ws);
    DPC_LAUNCH_CHECK();
    return DPC_OK;
}

// a = relu(bn1(ypool)) -> split-bf16 operand planes (and fp32 rows if a_rows != NULL); ReLU-dead windows get bit 7 of idx
extern "C" int dpc_stem_pool_finalize(const float* ypool, void* idx, const float* mean, const float* rstd, const float* gamma,
                                      const float* beta, void* a_hi, void* a_lo, float* a_rows, int64_t rows, void* stream) {
    DPC_REQUIRE(ypool && idx && mean && rstd && gamma && beta && a_hi && a_lo && rows > 0, "dpc_stem_pool_finalize: bad args");
    const long long n4 = rows * 16;
    long long blocks = (n4 + 255) / 256;
    const long long cap = (long long)dpc_num_sms() * 16;
    stem_pool_finalize_kernel<<<(int)(blocks < cap ? blocks : cap), 256, 0, as_stream(stream)>>>(
        reinterpret_cast<const float4*>(ypool), reinterpret_cast<uchar4*>(idx), mean, rstd, gamma, beta, a_hi, a_lo,
        reinterpret_cast<float4*>(a_rows), n4);
    DPC_LAUNCH_CHECK();
    return DPC_OK;
}

// bn1 backward sums from the pooled grid: ws [128 doubles] = sum g | sum g*xhat; dgamma / dbeta [64]
extern "C" int dpc_stem_pool_bwd_reduce(const float* ypool, const float* dout, const void* idx, const float* mean,
                                        const float* rstd, double* ws, float* dgamma, float* dbeta, int64_t rows, void* stream) {
    DPC_REQUIRE(ypool && dout && idx && mean && rstd && ws && dgamma && dbeta && rows > 0, "dpc_stem_pool_bwd_reduce: bad args");
    cudaStream_t st = as_stream(stream);
    const long long n4 = rows * 16;
    long long blocks = (n4 + 256 * 32 - 1) / (256 * 32);
    const long long cap = (long long)dpc_num_sms() * 8;
    DPC_CUDA(cudaMemsetAsync(ws, 0, sizeof(double) * 128, st));
    stem_pool_bwd_reduce_kernel<<<(int)(blocks < cap ? (blocks > 0 ? blocks : 1) : cap), 256, 0, st>>>(
        reinterpret_cast<const float4*>(ypool), reinterpret_cast<const float4*>(dout), reinterpret_cast<const uchar4*>(idx),
        mean, rstd, n4, ws);
    DPC_LAUNCH_CHECK();
    stem_pool_bwd_finalize_kernel<<<1, 64, 0, st>>>(ws, dgamma, dbeta);
    DPC_LAUNCH_CHECK();
    return DPC_OK;
}

namespace {
// fuse != 0: conv1's wgrad is issued by the same kernel (dw [64,3,1,7,7] accumulated with atomics, zeroed here);
// fuse == 0: the gradient planes dy_hi / dy_lo go to HBM
int sp_bwd_launch(int fuse, const void* x2_hi, const void* x2_lo, const void* wp, const float* dout, const void* idx,
                  const float* mean, const float* rstd, const float* gamma, const double* ws, void* dy_hi, void* dy_lo,
                  float* dw, int NB, int T, int H, int W, void* stream, const char* who) {
    SpGeom g;
    if (const char* e = sp_geometry(g, NB, T, H, W)) { dpc_set_error("%s: %s (H %d, W %d)", who, e, H, W); return DPC_ERR_UNSUPPORTED; }
    cudaStream_t st = as_stream(stream);
    if (fuse) {
        g.npb = 3;                 // the input patch of a tile lives until its wgrad MMAs retire: 3 buffers keep the conv fed
        if (sp_bwd_smem(g, true) > 227 * 1024) { dpc_set_error("%s: fused schedule does not fit shared memory (H %d, W %d)", who, H, W); return DPC_ERR_UNSUPPORTED; }
    } else {
        while (g.npb < SP_NPB_MAX) { ++g.npb; if (sp_bwd_smem(g, false) > 227 * 1024) { --g.npb; break; } }
    }
    SpMaps maps;
    memset(&maps, 0, sizeof(maps));
    if (int rc = sp_x_maps(maps, g, x2_hi, x2_lo, wp, NB)) return rc;
    auto enc = sp_encode();
    const cuuint32_t es[5] = {1, 1, 1, 1, 1};
    {
        const cuuint64_t gd[5] = {64, (cuuint64_t)g.Wp, (cuuint64_t)g.Hp, (cuuint64_t)g.T, (cuuint64_t)NB};
        const cuuint64_t gs[4] = {256, (cuuint64_t)g.Wp * 256, (cuuint64_t)g.Hp * g.Wp * 256, (cuuint64_t)g.T * g.Hp * g.Wp * 256};
        const cuuint32_t bx[5] = {32, (cuuint32_t)g.bc, (cuuint32_t)g.br, 1, 1};
        CUresult r = enc(&maps.dout, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, const_cast<float*>(dout), gd, gs, bx, es,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        DPC_REQUIRE(r == CUDA_SUCCESS, "%s: cuTensorMapEncodeTiled (dout) failed (%d)", who, (int)r);
        const cuuint64_t is[4] = {64, (cuuint64_t)g.Wp * 64, (cuuint64_t)g.Hp * g.Wp * 64, (cuuint64_t)g.T * g.Hp * g.Wp * 64};
        const cuuint32_t ib[5] = {64, (cuuint32_t)g.bc, (cuuint32_t)g.br, 1, 1};
        r = enc(&maps.didx, CU_TENSOR_MAP_DATA_TYPE_UINT8, 5, const_cast<void*>(idx), gd, is, ib, es,
                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        DPC_REQUIRE(r == CUDA_SUCCESS, "%s: cuTensorMapEncodeTiled (idx) failed (%d)", who, (int)r);
    }
    const size_t smem = sp_bwd_smem(g, fuse != 0);
    const int sms = dpc_num_sms();
    const int grid = g.total_units < sms ? g.total_units : sms;
    if (fuse) {
        DPC_CUDA(cudaMemsetAsync(dw, 0, sizeof(float) * 64 * 147, st));
        DPC_CUDA(cudaFuncSetAttribute(stem_pool_bwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        stem_pool_bwd_kernel<true><<<grid, SP_THREADS, smem, st>>>(maps, g, mean, rstd, gamma, ws, nullptr, nullptr, dw);
    } else {
        DPC_CUDA(cudaFuncSetAttribute(stem_pool_bwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        stem_pool_bwd_kernel<false><<<grid, SP_THREADS, smem, st>>>(maps, g, mean, rstd, gamma, ws, reinterpret_cast<uint16_t*>(dy_hi),
                                                                    reinterpret_cast<uint16_t*>(dy_lo), nullptr);
    }
    DPC_LAUNCH_CHECK();
    return DPC_OK;
}
}  // namespace

// gradient planes dy_hi / dy_lo [NB,T,H/2,W/2,64] of conv1's output from dout [NB,T,Hp,Wp,64] (gradient of the pooled,
// normalised output), idx (dpc_stem_pool_fwd + finalize) and ws (dpc_stem_pool_bwd_reduce); conv1 is recomputed.
extern "C" int dpc_stem_pool_bwd(const void* x2_hi, const void* x2_lo, const void* wp, const float* dout, const void* idx,
                                 const float* mean, const float* rstd, const float* gamma, const double* ws, void* dy_hi,
                                 void* dy_lo, int NB, int T, int H, int W, void* stream) {
    DPC_REQUIRE(x2_hi && x2_lo && wp && dout && idx && mean && rstd && gamma && ws && dy_hi && dy_lo && NB > 0 && T > 0,
                "dpc_stem_pool_bwd: bad args");
    return sp_bwd_launch(0, x2_hi, x2_lo, wp, dout, idx, mean, rstd, gamma, ws, dy_hi, dy_lo, nullptr, NB, T, H, W, stream,
                         "dpc_stem_pool_bwd");
}

// the same backward with conv1's weight gradient computed in the same kernel: dw [64,3,1,7,7] (overwritten); the gradient
// on the conv1 grid only ever exists as one 128-position tile in shared memory.  DPC_ERR_UNSUPPORTED when the fused
// schedule does not fit (dpc_stem_pool_supported(H, W) < 2): use dpc_stem_pool_bwd + dpc_stem_conv_wgrad_s2d then.
extern "C" int dpc_stem_pool_bwd_wgrad(const void* x2_hi, const void* x2_lo, const void* wp, const float* dout, const void* idx,
                                       const float* mean, const float* rstd, const float* gamma, const double* ws, float* dw,
                                       int NB, int T, int H, int W, void* stream) {
    DPC_REQUIRE(x2_hi && x2_lo && wp && dout && idx && mean && rstd && gamma && ws && dw && NB > 0 && T > 0,
                "dpc_stem_pool_bwd_wgrad: bad args");
    return sp_bwd_launch(1, x2_hi, x2_lo, wp, dout, idx, mean, rstd, gamma, ws, nullptr, nullptr, dw, NB, T, H, W, stream,
                         "dpc_stem_pool_bwd_wgrad");
}

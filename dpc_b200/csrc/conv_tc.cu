// tcgen05 (5th-gen tensor core) implicit-GEMM for the 1x3x3 / 3x3x3 / 1x1x1 convolutions and the
// dense score matmul, sm_100a only.
//
//   D[128 x BN] (fp32, TMEM) += sum over k-blocks  A_hi*B_hi + A_hi*B_lo + A_lo*B_hi      (3xBF16 split)
//
// * operands live in HBM as two bf16 planes (hi = bf16(x), lo = bf16(x - hi)): same bytes as fp32,
//   ~16 mantissa bits, which is what the 1e-3 parity bar needs (single-pass
//   BF16/TF32 fail it);
// * A tile (128 output positions x 64 channels of ONE filter tap) is ONE TMA box of the channels-last
//   activation tensor [NB,T,H,W,C] at the tap-shifted coordinate; the halo / zero padding is TMA
//   out-of-bounds fill, so there is no im2col buffer and no index arithmetic on the SM;
// * B tile (BN filters x 64 channels of that tap) is a TMA box of the packed weights [Co][tap][Ci];
// * both land in shared memory in the 128B-swizzled K-major layout the UMMA descriptors expect;
// * warp 0 = TMA producer, warp 1 = MMA issuer (one elected lane) + TMEM owner, warps 2-5 = epilogue
//   (tcgen05.ld -> registers -> coalesced fp32 row stores), mbarrier full/empty ring between them.
//
// Replaces nn.Conv3d at backbone/resnet_2d3d.py:13-31,241-244 (stride-1 sites; strided sites go
// through per-parity tensor maps, see dpc_conv3d_fwd_tc) and torch.matmul at dpc/model_3d.py:83.
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
// Bounded wait: a protocol bug traps (-> CUDA error on the host) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    const long long t0 = clock64();
    for (;;) {
        asm volatile(
            "{\n\t"
            ".reg .pred P1;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, P1;\n\t"
            "}" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
        if (done) return;
        if (clock64() - t0 > 4000000000ll) break;          // ~2 s: far beyond any legitimate wait
    }
    printf("dpc_b200: mbarrier wait timed out (block %d,%d thread %d bar 0x%x parity %u)\n", blockIdx.x, blockIdx.y,
           threadIdx.x, bar, parity);
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
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
          "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
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

struct TcParams {
    // filter taps: k-block kb = tap * cchunks + cc
    int taps, kH, kW, cchunks;
    int offT, offH, offW;      // coordinate of tap (0,0,0) relative to the output position (= -pad)
    int Ksrc;                  // channels of A (= cchunks * 64)
    // A box (output-position tile)
    int bw, bh, bt, bn, box_rows;
    int tiles_w, tiles_h, tiles_t;
    // output tensor extents and channel count
    int NB, To, Ho, Wo, Co;
    int BN, stages;
    long long out_sn, out_st, out_sh, out_sw;    // output row index strides (rows), for strided dgrad
    int out_t0, out_h0, out_w0;                  // output origin (parity class), rows
};

constexpr int A_TILE_BYTES = 128 * 128;            // 128 rows x 64 bf16

__global__ void __launch_bounds__(192, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap mAhi, const __grid_constant__ CUtensorMap mAlo,
               const __grid_constant__ CUtensorMap mBhi, const __grid_constant__ CUtensorMap mBlo,
               const TcParams p, float* __restrict__ y, int accumulate) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // carve: [stage][Ahi | Alo | Bhi | Blo], then barriers
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const int b_tile_bytes = p.BN * 128;
    const int stage_bytes = 2 * A_TILE_BYTES + 2 * b_tile_bytes;
    const uint32_t bar_base = smem_base + p.stages * stage_bytes;      // full[s], empty[s], tmem_full, tmem_ptr
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (p.stages + s); };
    const uint32_t tmem_full_bar = bar_base + 8u * (2 * p.stages);
    const uint32_t tmem_ptr_addr = bar_base + 8u * (2 * p.stages + 1);
    uint32_t* tmem_ptr_gen = reinterpret_cast<uint32_t*>(smem_raw + (tmem_ptr_addr - smem_u32(smem_raw)));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t tmem_cols = p.BN < 32 ? 32 : p.BN;     // power of two >= 32 (BN in {32,64,128,256})

    // tile coordinates
    int tile = blockIdx.x;
    const int tw = tile % p.tiles_w; tile /= p.tiles_w;
    const int th = tile % p.tiles_h; tile /= p.tiles_h;
    const int tt = tile % p.tiles_t; tile /= p.tiles_t;
    const int tn = tile;
    const int w0 = tw * p.bw, h0 = th * p.bh, t0 = tt * p.bt, n0 = tn * p.bn;
    const int ncol0 = blockIdx.y * p.BN;
    const int num_kb = p.taps * p.cchunks;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&mAhi) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&mAlo) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&mBhi) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&mBlo) : "memory");
        for (int s = 0; s < p.stages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
        mbar_init(tmem_full_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) tmem_alloc(tmem_ptr_addr, tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_d = *tmem_ptr_gen;

    if (warp == 0) {
        // ===== TMA producer =====
        if (elect_one()) {
            const uint32_t tx = 2u * (uint32_t)(p.box_rows * 128) + 2u * (uint32_t)b_tile_bytes;
            int s = 0; uint32_t ph = 0;
            for (int kb = 0; kb < num_kb; ++kb) {
                const int tap = kb / p.cchunks, cc = kb - tap * p.cchunks;
                const int kt = tap / (p.kH * p.kW), kh = (tap / p.kW) % p.kH, kw = tap % p.kW;
                mbar_wait(empty_bar(s), ph ^ 1u);
                mbar_expect_tx(full_bar(s), tx);
                const uint32_t sa = smem_base + s * stage_bytes;
                const int cw = w0 + kw + p.offW, chh = h0 + kh + p.offH, ct = t0 + kt + p.offT;
                tma_load_5d(&mAhi, sa, full_bar(s), cc * 64, cw, chh, ct, n0);
                tma_load_5d(&mAlo, sa + A_TILE_BYTES, full_bar(s), cc * 64, cw, chh, ct, n0);
                const int kcol = tap * p.Ksrc + cc * 64;
                tma_load_2d(&mBhi, sa + 2 * A_TILE_BYTES, full_bar(s), kcol, ncol0);
                tma_load_2d(&mBlo, sa + 2 * A_TILE_BYTES + b_tile_bytes, full_bar(s), kcol, ncol0);
                if (++s == p.stages) { s = 0; ph ^= 1u; }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        if (elect_one()) {
            // instruction descriptor: D=f32, A=B=bf16, both K-major, N = BN, M = 128
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.BN >> 3) << 17) | ((128u >> 4) << 24);
            int s = 0; uint32_t ph = 0;
            for (int kb = 0; kb < num_kb; ++kb) {
                mbar_wait(full_bar(s), ph);
                tc_fence_after();
                const uint32_t sa = smem_base + s * stage_bytes;
                const uint64_t ahi = make_kmajor_sw128_desc(sa), alo = make_kmajor_sw128_desc(sa + A_TILE_BYTES);
                const uint64_t bhi = make_kmajor_sw128_desc(sa + 2 * A_TILE_BYTES);
                const uint64_t blo = make_kmajor_sw128_desc(sa + 2 * A_TILE_BYTES + b_tile_bytes);
#pragma unroll
                for (int k = 0; k < 4; ++k) {               // 4 x UMMA_K(16) = 64 channels; +32 B per step
                    const uint64_t ko = (uint64_t)(k * 2);
                    umma_bf16(tmem_d, ahi + ko, bhi + ko, idesc, (kb | k) ? 1u : 0u);
                    umma_bf16(tmem_d, ahi + ko, blo + ko, idesc, 1u);
                    umma_bf16(tmem_d, alo + ko, bhi + ko, idesc, 1u);
                }
                umma_commit(empty_bar(s));                  // frees the smem stage when these MMAs retire
                if (++s == p.stages) { s = 0; ph ^= 1u; }
            }
            umma_commit(tmem_full_bar);                     // accumulator complete
        }
    } else {
        // ===== epilogue: warps 2..5; TMEM lane quarter = warp % 4 =====
        const int q = warp & 3;
        const int r = q * 32 + lane;                        // tile row == TMEM lane
        const int dw = r % p.bw, dh = (r / p.bw) % p.bh, dt = (r / (p.bw * p.bh)) % p.bt, dn = r / (p.bw * p.bh * p.bt);
        const int n = n0 + dn, t = t0 + dt, h = h0 + dh, w = w0 + dw;
        const bool valid = r < p.box_rows && n < p.NB && t < p.To && h < p.Ho && w < p.Wo;
        const long long row = (long long)n * p.out_sn + (long long)(t + p.out_t0) * p.out_st +
                              (long long)(h + p.out_h0) * p.out_sh + (long long)(w + p.out_w0) * p.out_sw;
        float* yrow = y + row * p.Co + ncol0;
        mbar_wait(tmem_full_bar, 0);
        tc_fence_after();
        const bool vec = (p.Co & 3) == 0;
        for (int c0 = 0; c0 < p.BN; c0 += 32) {
            uint32_t v[32];
            tmem_ld32(tmem_d + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
            if (!valid) continue;
            if (vec && ncol0 + c0 + 32 <= p.Co) {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    float4 o = make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]),
                                           __uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3]));
                    float4* dst = reinterpret_cast<float4*>(yrow + c0) + j;
                    if (accumulate) { float4 c = *dst; o.x += c.x; o.y += c.y; o.z += c.z; o.w += c.w; }
                    *dst = o;
                }
            } else {
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    if (ncol0 + c0 + j < p.Co) {
                        float o = __uint_as_float(v[j]);
                        if (accumulate) o += yrow[c0 + j];
                        yrow[c0 + j] = o;
                    }
                }
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

// bf16 tensor [d4][d3][d2][d1][d0] (d0 contiguous) with arbitrary byte strides for d1..d4
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
    DPC_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed (%d) rank %d dims %llu %llu box %u %u", (int)r, rank,
                (unsigned long long)dims[0], (unsigned long long)dims[1], box[0], box[1]);
    return DPC_OK;
}

// pick the box (bw,bh,bt,bn), product <= 128, maximising coverage efficiency per dimension
// Every tile costs a full 128-row MMA, so minimise the tile count; ties -> longer contiguous rows.
void choose_box(int W, int H, int T, int NB, int& bw, int& bh, int& bt, int& bn) {
    auto cdiv = [](long long a, long long b) { return (a + b - 1) / b; };
    long long best = -1;
    bw = bh = bt = bn = 1;
    for (int w = (W < 128 ? W : 128); w >= 1; --w) {
        const int rh = 128 / w;
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
    CUtensorMap mAhi, mAlo, mBhi, mBlo;
    dim3 grid;
    size_t smem;
};

// A: bf16 planes of a channels-last tensor with extents (NB, Ta, Ha, Wa, Ca) and row strides
// (in elements) sn, st, sh, sw (sw = Ca for a dense tensor; parity views use multiples).
int setup(TcLaunch& L, const void* a_hi, const void* a_lo, int NB, int Ta, int Ha, int Wa, int Ca,
          long long sn, long long st, long long sh, long long sw,
          const void* b_hi, const void* b_lo, int Co, int Ktot,
          int To, int Ho, int Wo, int taps, int kH, int kW, int offT, int offH, int offW) {
    TcParams& p = L.p;
    DPC_REQUIRE(Ca % 64 == 0, "tcgen05 conv: channel count %d must be a multiple of 64", Ca);
    p.taps = taps; p.kH = kH; p.kW = kW; p.cchunks = Ca / 64; p.Ksrc = Ca;
    p.offT = offT; p.offH = offH; p.offW = offW;
    choose_box(Wo, Ho, To, NB, p.bw, p.bh, p.bt, p.bn);
    p.box_rows = p.bw * p.bh * p.bt * p.bn;
    p.tiles_w = (Wo + p.bw - 1) / p.bw; p.tiles_h = (Ho + p.bh - 1) / p.bh; p.tiles_t = (To + p.bt - 1) / p.bt;
    const int tiles_n = (NB + p.bn - 1) / p.bn;
    p.NB = NB; p.To = To; p.Ho = Ho; p.Wo = Wo; p.Co = Co;
    p.BN = Co >= 256 ? 256 : (Co >= 128 ? 128 : (Co >= 64 ? 64 : 32));
    const int stage_bytes = 2 * A_TILE_BYTES + 2 * p.BN * 128;
    int stages = (200 * 1024) / stage_bytes;
    if (stages > 6) stages = 6;
    const int num_kb = taps * p.cchunks;
    if (stages > num_kb) stages = num_kb;
    DPC_REQUIRE(stages >= 1, "tcgen05 conv: no pipeline stage fits");
    p.stages = stages;
    L.smem = (size_t)stages * stage_bytes + 8 * (2 * stages + 2) + 1024;
    // default: dense output [NB,To,Ho,Wo]
    p.out_sw = 1; p.out_sh = Wo; p.out_st = (long long)Ho * Wo; p.out_sn = (long long)To * Ho * Wo;
    p.out_t0 = p.out_h0 = p.out_w0 = 0;
    const uint64_t ad[5] = {(uint64_t)Ca, (uint64_t)Wa, (uint64_t)Ha, (uint64_t)Ta, (uint64_t)NB};
    const uint64_t as[4] = {(uint64_t)sw * 2, (uint64_t)sh * 2, (uint64_t)st * 2, (uint64_t)sn * 2};
    const uint32_t ab[5] = {64, (uint32_t)p.bw, (uint32_t)p.bh, (uint32_t)p.bt, (uint32_t)p.bn};
    if (int rc = make_map(&L.mAhi, a_hi, 5, ad, as, ab)) return rc;
    if (int rc = make_map(&L.mAlo, a_lo, 5, ad, as, ab)) return rc;
    const uint64_t bd[2] = {(uint64_t)Ktot, (uint64_t)Co};
    const uint64_t bs[1] = {(uint64_t)Ktot * 2};
    const uint32_t bb[2] = {64, (uint32_t)p.BN};
    if (int rc = make_map(&L.mBhi, b_hi, 2, bd, bs, bb)) return rc;
    if (int rc = make_map(&L.mBlo, b_lo, 2, bd, bs, bb)) return rc;
    L.grid = dim3((unsigned)(p.tiles_w * p.tiles_h * p.tiles_t * tiles_n), (unsigned)((Co + p.BN - 1) / p.BN));
    return DPC_OK;
}

int launch(TcLaunch& L, float* y, int accumulate, cudaStream_t st) {
    DPC_CUDA(cudaFuncSetAttribute(conv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.smem));
    conv_tc_kernel<<<L.grid, 192, L.smem, st>>>(L.mAhi, L.mAlo, L.mBhi, L.mBlo, L.p, y, accumulate);
    DPC_LAUNCH_CHECK();
    return DPC_OK;
}

__global__ void split_bf16_kernel(const float4* __restrict__ src, uint2* __restrict__ hi, uint2* __restrict__ lo,
                                  long long n4) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        float4 v = src[i];
        float f[4] = {v.x, v.y, v.z, v.w};
        unsigned short h[4], l[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            __nv_bfloat16 bh = __float2bfloat16_rn(f[j]);
            __nv_bfloat16 bl = __float2bfloat16_rn(f[j] - __bfloat162float(bh));
            h[j] = __bfloat16_as_ushort(bh);
            l[j] = __bfloat16_as_ushort(bl);
        }
        hi[i] = make_uint2((uint32_t)h[0] | ((uint32_t)h[1] << 16), (uint32_t)h[2] | ((uint32_t)h[3] << 16));
        lo[i] = make_uint2((uint32_t)l[0] | ((uint32_t)l[1] << 16), (uint32_t)l[2] | ((uint32_t)l[3] << 16));
    }
}

// w [Co][Ci][taps] fp32 -> forward planes [Co][tap][Ci] and dgrad planes [Ci][tap'][Co] (tap' = flipped)
__global__ void pack_weight_bf16_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ fh,
                                        __nv_bfloat16* __restrict__ fl, __nv_bfloat16* __restrict__ dh,
                                        __nv_bfloat16* __restrict__ dl, int Co, int Ci, int taps) {
    long long total = (long long)Co * Ci * taps;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        // i indexes the forward layout [co][tap][ci]
        int ci = (int)(i % Ci);
        int tap = (int)((i / Ci) % taps);
        int co = (int)(i / ((long long)Ci * taps));
        float v = w[((size_t)co * Ci + ci) * taps + tap];
        __nv_bfloat16 h = __float2bfloat16_rn(v);
        __nv_bfloat16 l = __float2bfloat16_rn(v - __bfloat162float(h));
        if (fh) { fh[i] = h; fl[i] = l; }
        if (dh) {
            size_t j = ((size_t)ci * taps + (taps - 1 - tap)) * Co + co;
            dh[j] = h; dl[j] = l;
        }
    }
}

}  // namespace

extern "C" int dpc_split_bf16(const float* src, void* hi, void* lo, int64_t n, void* stream) {
    DPC_REQUIRE(src && hi && lo && n > 0 && n % 4 == 0, "dpc_split_bf16: bad args (n must be a multiple of 4)");
    long long n4 = n / 4;
    long long blocks = (n4 + 255) / 256;
    long long cap = (long long)dpc_num_sms() * 16;
    split_bf16_kernel<<<(int)(blocks < cap ? blocks : cap), 256, 0, as_stream(stream)>>>(
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

// C[M,N] (fp32, ldc = N) = A[M,K] * B[N,K]^T from split-bf16 planes (K % 64 == 0)
extern "C" int dpc_gemm_nt_bf16x3_tc(int M, int N, int K, const void* a_hi, const void* a_lo, const void* b_hi,
                                     const void* b_lo, float* C, void* stream) {
    DPC_REQUIRE(M > 0 && N > 0 && K > 0 && K % 64 == 0, "dpc_gemm_nt_bf16x3_tc: bad dims %d %d %d", M, N, K);
    DPC_REQUIRE(a_hi && a_lo && b_hi && b_lo && C, "dpc_gemm_nt_bf16x3_tc: null pointer");
    TcLaunch L;
    // a [1,1,1,M,K] "image" convolved with a 1x1x1 filter bank of N filters
    if (int rc = setup(L, a_hi, a_lo, 1, 1, 1, M, K, (long long)M * K, (long long)M * K, (long long)M * K, K, b_hi, b_lo,
                       N, K, 1, 1, M, 1, 1, 1, 0, 0, 0))
        return rc;
    return launch(L, C, 0, as_stream(stream));
}

// stride-1 convolution (forward, or dgrad when called with dy planes and flipped/transposed weights):
// y [NB,To,Ho,Wo,Co] (+)= conv(x planes [NB,Ti,Hi,Wi,Ci], w planes [Co][taps][Ci])
extern "C" int dpc_conv3d_s1_tc(const dpc_conv_geom* g, const void* x_hi, const void* x_lo, const void* w_hi,
                                const void* w_lo, float* y, int accumulate, void* stream) {
    DPC_REQUIRE(g && x_hi && x_lo && w_hi && w_lo && y, "dpc_conv3d_s1_tc: null pointer");
    DPC_REQUIRE(g->sT == 1 && g->sH == 1 && g->sW == 1, "dpc_conv3d_s1_tc: stride must be 1");
    DPC_REQUIRE((g->Ti + 2 * g->pT - g->kT) + 1 == g->To && (g->Hi + 2 * g->pH - g->kH) + 1 == g->Ho &&
                    (g->Wi + 2 * g->pW - g->kW) + 1 == g->Wo,
                "dpc_conv3d_s1_tc: output extent does not match the conv arithmetic");
    const int taps = g->kT * g->kH * g->kW;
    TcLaunch L;
    const long long sw = g->Ci, sh = (long long)g->Wi * sw, st = (long long)g->Hi * sh, sn = (long long)g->Ti * st;
    if (int rc = setup(L, x_hi, x_lo, g->NB, g->Ti, g->Hi, g->Wi, g->Ci, sn, st, sh, sw, w_hi, w_lo, g->Co,
                       taps * g->Ci, g->To, g->Ho, g->Wo, taps, g->kH, g->kW, -g->pT, -g->pH, -g->pW))
        return rc;
    return launch(L, y, accumulate, as_stream(stream));
}

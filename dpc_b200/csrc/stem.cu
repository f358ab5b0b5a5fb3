// Stem: Conv3d(3,64,(1,7,7),stride(1,2,2),pad(0,3,3)) reading the caller's NCDHW fp32 input directly
// (fuses the layout change) and writing channels-last rows.  HBM/CUDA-core kernel by design
// (Cin = 3, K = 147: not tensor-core shaped).  Replaces backbone/resnet_2d3d.py:211,260.
#include "common.cuh"

namespace {

constexpr int TH = 8, TW = 32;                 // output tile (pixels)
constexpr int PH = 2 * TH + 5, PW = 2 * TW + 5; // input patch 21 x 69
constexpr int KK = 147;                        // 3*7*7
constexpr int PATCH = 3 * PH * PW;             // 4347 floats

__device__ __forceinline__ void load_patch(float* patch, const float* __restrict__ x, int n, int t,
                                           int T, int H, int W, int ho0, int wo0) {
    for (int i = threadIdx.x; i < PATCH; i += blockDim.x) {
        int col = i % PW, r = (i / PW) % PH, c = i / (PW * PH);
        int hi = 2 * ho0 - 3 + r, wi = 2 * wo0 - 3 + col;
        float v = 0.f;
        if (hi >= 0 && hi < H && wi >= 0 && wi < W)
            v = x[((((size_t)n * 3 + c) * T + t) * H + hi) * W + wi];
        patch[i] = v;
    }
}

__global__ void __launch_bounds__(256) stem_fwd_kernel(const float* __restrict__ x,
                                                        const float* __restrict__ w,
                                                        float* __restrict__ y, int NB, int T, int H,
                                                        int W, int Ho, int Wo) {
    extern __shared__ __align__(16) float smem[];
    float* ws = smem;                  // [147][64]
    float* patch = smem + KK * 64;     // [3][21][69]
    const int tid = threadIdx.x;
    for (int i = tid; i < KK * 64; i += 256) {
        int co = i % 64, k = i / 64;
        ws[i] = w[co * KK + k];
    }
    const int tiles_w = (Wo + TW - 1) / TW, tiles_h = (Ho + TH - 1) / TH;
    int b = blockIdx.x;
    const int tw = b % tiles_w; b /= tiles_w;
    const int th = b % tiles_h; b /= tiles_h;
    const int t = b % T;
    const int n = b / T;
    const int ho0 = th * TH, wo0 = tw * TW;
    load_patch(patch, x, n, t, T, H, W, ho0, wo0);
    __syncthreads();

    const int half = tid / 128, q = tid % 128;
    const int py = q / 16, px = q % 16;
    float acc0[32], acc1[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) { acc0[j] = 0.f; acc1[j] = 0.f; }
    const float* wbase = ws + half * 32;
    for (int c = 0; c < 3; ++c) {
        for (int kh = 0; kh < 7; ++kh) {
            const float* prow = patch + (c * PH + 2 * py + kh) * PW + 2 * px;
#pragma unroll
            for (int kw = 0; kw < 7; ++kw) {
                float a0 = prow[kw], a1 = prow[kw + 32];
                const float4* wv = reinterpret_cast<const float4*>(wbase + ((c * 7 + kh) * 7 + kw) * 64);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    float4 f = wv[j];
                    acc0[4 * j + 0] = fmaf(a0, f.x, acc0[4 * j + 0]);
                    acc0[4 * j + 1] = fmaf(a0, f.y, acc0[4 * j + 1]);
                    acc0[4 * j + 2] = fmaf(a0, f.z, acc0[4 * j + 2]);
                    acc0[4 * j + 3] = fmaf(a0, f.w, acc0[4 * j + 3]);
                    acc1[4 * j + 0] = fmaf(a1, f.x, acc1[4 * j + 0]);
                    acc1[4 * j + 1] = fmaf(a1, f.y, acc1[4 * j + 1]);
                    acc1[4 * j + 2] = fmaf(a1, f.z, acc1[4 * j + 2]);
                    acc1[4 * j + 3] = fmaf(a1, f.w, acc1[4 * j + 3]);
                }
            }
        }
    }
    const int ho = ho0 + py;
    if (ho < Ho) {
        size_t rowbase = (((size_t)n * T + t) * Ho + ho) * Wo;
        int wo = wo0 + px;
        if (wo < Wo) {
            float4* o = reinterpret_cast<float4*>(y + (rowbase + wo) * 64 + half * 32);
#pragma unroll
            for (int j = 0; j < 8; ++j) o[j] = make_float4(acc0[4 * j], acc0[4 * j + 1], acc0[4 * j + 2], acc0[4 * j + 3]);
        }
        wo += 16;
        if (wo < Wo) {
            float4* o = reinterpret_cast<float4*>(y + (rowbase + wo) * 64 + half * 32);
#pragma unroll
            for (int j = 0; j < 8; ++j) o[j] = make_float4(acc1[4 * j], acc1[4 * j + 1], acc1[4 * j + 2], acc1[4 * j + 3]);
        }
    }
}

// wgrad: persistent CTAs loop over pixel tiles; thread (kg, cg) owns k = kg + 16*j (j < 10), 4 co.
constexpr int KJ = 10;
__global__ void __launch_bounds__(256) stem_wgrad_kernel(const float* __restrict__ x,
                                                          const float* __restrict__ dy,
                                                          float* __restrict__ dw, int NB, int T, int H,
                                                          int W, int Ho, int Wo, int total_tiles) {
    extern __shared__ __align__(16) float smem[];
    float* patch = smem;                       // [3][21][69]  (4347, padded to 4352)
    float* dys = smem + 4352;                  // [256 pix][64]
    const int tid = threadIdx.x;
    const int cg = tid % 16, kg = tid / 16;
    int koff[KJ];
    bool kval[KJ];
#pragma unroll
    for (int j = 0; j < KJ; ++j) {
        int k = kg + 16 * j;
        kval[j] = k < KK;
        int kk = kval[j] ? k : 0;
        int c = kk / 49, kh = (kk / 7) % 7, kw = kk % 7;
        koff[j] = (c * PH + kh) * PW + kw;
    }
    float acc[KJ][4];
#pragma unroll
    for (int j = 0; j < KJ; ++j)
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[j][e] = 0.f;

    const int tiles_w = (Wo + TW - 1) / TW, tiles_h = (Ho + TH - 1) / TH;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        int b = tile;
        const int tw = b % tiles_w; b /= tiles_w;
        const int th = b % tiles_h; b /= tiles_h;
        const int t = b % T;
        const int n = b / T;
        const int ho0 = th * TH, wo0 = tw * TW;
        __syncthreads();
        load_patch(patch, x, n, t, T, H, W, ho0, wo0);
        for (int i = tid; i < 256 * 16; i += 256) {
            int pix = i / 16, c4 = i % 16;
            int ho = ho0 + pix / TW, wo = wo0 + pix % TW;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (ho < Ho && wo < Wo)
                v = *reinterpret_cast<const float4*>(dy + ((((size_t)n * T + t) * Ho + ho) * Wo + wo) * 64 + c4 * 4);
            *reinterpret_cast<float4*>(dys + pix * 64 + c4 * 4) = v;
        }
        __syncthreads();
        for (int pix = 0; pix < 256; ++pix) {
            const int py = pix / TW, px = pix % TW;
            const float4 d = *reinterpret_cast<const float4*>(dys + pix * 64 + cg * 4);
            const float* pp = patch + (2 * py) * PW + 2 * px;
#pragma unroll
            for (int j = 0; j < KJ; ++j) {
                float a = pp[koff[j]];
                acc[j][0] = fmaf(a, d.x, acc[j][0]);
                acc[j][1] = fmaf(a, d.y, acc[j][1]);
                acc[j][2] = fmaf(a, d.z, acc[j][2]);
                acc[j][3] = fmaf(a, d.w, acc[j][3]);
            }
        }
    }
#pragma unroll
    for (int j = 0; j < KJ; ++j) {
        if (!kval[j]) continue;
        int k = kg + 16 * j;
#pragma unroll
        for (int e = 0; e < 4; ++e) atomicAdd(dw + (size_t)(cg * 4 + e) * KK + k, acc[j][e]);
    }
}

}  // namespace

extern "C" int dpc_stem_conv_fwd(const float* x, const float* w, float* y, int NB, int T, int H, int W, void* stream) {
    DPC_REQUIRE(x && w && y && NB > 0 && T > 0 && H > 0 && W > 0, "dpc_stem_conv_fwd: bad args");
    const int Ho = (H + 6 - 7) / 2 + 1, Wo = (W + 6 - 7) / 2 + 1;
    const int tiles = NB * T * ((Ho + TH - 1) / TH) * ((Wo + TW - 1) / TW);
    const size_t smem = (size_t)(KK * 64 + PATCH + 4) * sizeof(float);
    DPC_CUDA(cudaFuncSetAttribute(stem_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    stem_fwd_kernel<<<tiles, 256, smem, as_stream(stream)>>>(x, w, y, NB, T, H, W, Ho, Wo);
    DPC_LAUNCH_CHECK();
    return DPC_OK;
}

extern "C" int dpc_stem_conv_wgrad(const float* x, const float* dy, float* dw, int NB, int T, int H, int W, void* stream) {
    DPC_REQUIRE(x && dy && dw && NB > 0 && T > 0 && H > 0 && W > 0, "dpc_stem_conv_wgrad: bad args");
    const int Ho = (H + 6 - 7) / 2 + 1, Wo = (W + 6 - 7) / 2 + 1;
    const int tiles = NB * T * ((Ho + TH - 1) / TH) * ((Wo + TW - 1) / TW);
    const size_t smem = (size_t)(4352 + 256 * 64) * sizeof(float);
    DPC_CUDA(cudaFuncSetAttribute(stem_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaStream_t st = as_stream(stream);
    DPC_CUDA(cudaMemsetAsync(dw, 0, sizeof(float) * 64 * KK, st));
    int grid = dpc_num_sms() * 2;
    if (grid > tiles) grid = tiles;
    stem_wgrad_kernel<<<grid, 256, smem, st>>>(x, dy, dw, NB, T, H, W, Ho, Wo, tiles);
    DPC_LAUNCH_CHECK();
    return DPC_OK;
}

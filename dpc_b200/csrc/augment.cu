// On-device clip augmentation: decoded uint8 frames -> the float32 block DPC_RNN.forward consumes.
//
// Replaces the CPU data-loader transform chain of the reference,
//   /root/reference/utils/augmentation.py   RandomSizedCrop :147-203, RandomCrop :98-144, Scale :20-41,
//       RandomHorizontalFlip :206-232, RandomGray :235-261, ColorJitter :264-355, ToTensor :373-376, Normalize :378-384
//   composed as /root/reference/dpc/main.py:115-133 and reshaped as /root/reference/dpc/dataset_3d.py:108-112,
// including the Pillow / torchvision arithmetic those classes delegate to (Image.resize BILINEAR with the down-scale
// dependent support and 22-bit fixed-point coefficients, NEAREST, Image.blend in float32, convert("L"), the HSV round trip,
// ImageEnhance.Brightness / Contrast / Color, to_tensor + normalize), bit for bit: every intermediate image is uint8 exactly as
// in Pillow, float steps use non-contracted IEEE operations in the precision the C sources use.
//
// The host (dpc_b200/augmentation.py) draws every random decision in the reference's order and hands over, per clip, the
// separable resampling tables in SOURCE coordinates (crop, flips and resize are all folded into them) and, per frame, the
// grey channel and the ordered colour operations.  One CTA per frame:
//   1. resample: horizontal pass value per needed source row (uint8), then the vertical pass -- into a planar uint8 image
//      in shared memory (3 x Ho x Wo <= 147 KB at 224^2);
//   2. RandomGray channel replication; the ColorJitter operations in their per-frame order (the contrast step needs the
//      frame's mean luma: one block reduction);
//   3. ToTensor + Normalize, written straight into block[b, n, c, t, y, x] (coalesced along x).
// HBM-bound byte work: ~0.3 GB of frames in, 1.0 GB of block out per 128-clip step.
#include "common.cuh"

namespace {

constexpr int AUG_THREADS = 512;
constexpr int AUG_PREC = 22;                       // Resample.c PRECISION_BITS for 8-bit images

struct AugParams {
    const uint8_t* frames;                         // [B, F, H, W, 3]
    const int32_t* tables;                         // per clip: xstart[Wo] xcount[Wo] xcoef[Wo*K] ystart[Ho] ycount[Ho] ycoef[Ho*K] xstep
    const int32_t* fparams;                        // per frame: gray, op[4], factor bits[4], hue byte  (10 ints)
    float* out;                                    // [B, N, 3, SL, Ho, Wo]
    int B, F, H, W, Ho, Wo, K, N, SL;
    float mean[3], stdv[3];
};

__device__ __forceinline__ int clip8(int v) { return v < 0 ? 0 : (v > 255 ? 255 : v); }

// Image.blend(a, b, alpha) of Pillow's Blend.c: float32 a + alpha * (b - a) without contraction, truncated
__device__ __forceinline__ int blend8(int a, int b, float alpha, bool interp) {
    const float t = __fadd_rn((float)a, __fmul_rn(alpha, (float)(b - a)));
    if (interp) return (int)t & 255;
    return t <= 0.f ? 0 : (t >= 255.f ? 255 : (int)t);
}

__device__ __forceinline__ int luma(int r, int g, int b) { return (r * 19595 + g * 38470 + b * 7471 + 0x8000) >> 16; }

// convert("HSV"), hue byte += shift (mod 256), convert("RGB") -- Convert.c rgb2hsv_row / hsv2rgb
// `lut_i` / `lut_f` / `lut_fs`: per byte, floor(h * 6 / 255), its float remainder and (float)(s / 255) of hsv2rgb
__device__ __forceinline__ void hue_rotate(int& r, int& g, int& b, int shift, const int* lut_i, const float* lut_f,
                                           const float* lut_fs) {
    const int maxc = max(r, max(g, b)), minc = min(r, min(g, b));
    int uh = 0, us = 0;
    if (minc != maxc) {
        const float cr = (float)(maxc - minc);
        const float s = __fdiv_rn(cr, (float)maxc);
        const float rc = __fdiv_rn((float)(maxc - r), cr), gc = __fdiv_rn((float)(maxc - g), cr),
                    bc = __fdiv_rn((float)(maxc - b), cr);
        float h;
        if (r == maxc) h = __fsub_rn(bc, gc);
        else if (g == maxc) h = (float)__dsub_rn(__dadd_rn(2.0, (double)rc), (double)bc);
        else h = (float)__dsub_rn(__dadd_rn(4.0, (double)gc), (double)rc);
        h = (float)fmod(__dadd_rn(__ddiv_rn((double)h, 6.0), 1.0), 1.0);
        uh = clip8((int)__dmul_rn((double)h, 255.0));
        us = clip8((int)__dmul_rn((double)s, 255.0));
    }
    uh = (uh + shift) & 255;
    const int v = maxc;
    if (us == 0) { r = g = b = v; return; }
    const int i = lut_i[uh];
    const double f = (double)lut_f[uh];
    const double fs = (double)lut_fs[us];
    const double vf = (double)v;
    // C round() of a non-negative value == floor(x + 0.5) here
    const int p = clip8((int)floor(__dadd_rn(__dmul_rn(vf, __dsub_rn(1.0, fs)), 0.5)));
    const int q = clip8((int)floor(__dadd_rn(__dmul_rn(vf, __dsub_rn(1.0, __dmul_rn(fs, f))), 0.5)));
    const int t = clip8((int)floor(__dadd_rn(__dmul_rn(vf, __dsub_rn(1.0, __dmul_rn(fs, __dsub_rn(1.0, f)))), 0.5)));
    switch (i % 6) {
        case 0: r = v; g = t; b = p; break;
        case 1: r = q; g = v; b = p; break;
        case 2: r = p; g = v; b = t; break;
        case 3: r = p; g = q; b = v; break;
        case 4: r = t; g = p; b = v; break;
        default: r = v; g = p; b = q; break;
    }
}

__global__ void __launch_bounds__(AUG_THREADS)
augment_kernel(const AugParams p) {
    extern __shared__ uint8_t img[];                // [3][Ho * Wo]
    __shared__ unsigned long long red[AUG_THREADS / 32];
    __shared__ int s_mean;
    __shared__ int lut_i[256];
    __shared__ float lut_f[256], lut_fs[256], lut_out[3][256];
    if (threadIdx.x < 256) {
        const int u = threadIdx.x;
        const double hf = __ddiv_rn(__dmul_rn((double)u, 6.0), 255.0);
        const int i = (int)floor(hf);
        lut_i[u] = i;
        lut_f[u] = (float)__dsub_rn(hf, (double)i);
        lut_fs[u] = (float)__ddiv_rn((double)u, 255.0);
#pragma unroll
        for (int c = 0; c < 3; ++c)                     // ToTensor + Normalize of a byte, IEEE float32
            lut_out[c][u] = __fdiv_rn(__fsub_rn(__fdiv_rn((float)u, 255.f), p.mean[c]), p.stdv[c]);
    }
    __syncthreads();
    const int f = blockIdx.x % p.F, b = blockIdx.x / p.F;
    const int npix = p.Ho * p.Wo;
    const int tab_len = (p.Wo + p.Ho) * (2 + p.K) + 1;
    const int32_t* tab = p.tables + (size_t)b * tab_len;
    const int32_t *xstart = tab, *xcount = tab + p.Wo, *xcoef = tab + 2 * p.Wo;
    const int32_t *ystart = xcoef + p.Wo * p.K, *ycount = ystart + p.Ho, *ycoef = ystart + 2 * p.Ho;
    const int xstep = ycoef[p.Ho * p.K];
    const int32_t* fp = p.fparams + ((size_t)b * p.F + f) * 10;
    const int gray = fp[0];
    const uint8_t* src = p.frames + ((size_t)b * p.F + f) * p.H * p.W * 3;
    uint8_t *pr = img, *pg = img + npix, *pb = img + 2 * npix;

    // 1. separable resampling, horizontal pass first (uint8 after each pass, as Resample.c)
    for (int i = threadIdx.x; i < npix; i += AUG_THREADS) {
        const int y = i / p.Wo, x = i - y * p.Wo;
        const int x0 = xstart[x], xn = xcount[x], y0 = ystart[y], yn = ycount[y];
        const int32_t* kx = xcoef + x * p.K;
        const int32_t* ky = ycoef + y * p.K;
        int a0 = 1 << (AUG_PREC - 1), a1 = a0, a2 = a0;
        for (int j = 0; j < yn; ++j) {
            const uint8_t* row = src + (size_t)(y0 + j) * p.W * 3;
            int h0 = 1 << (AUG_PREC - 1), h1 = h0, h2 = h0;
            for (int k = 0; k < xn; ++k) {
                const uint8_t* px = row + (x0 + xstep * k) * 3;
                const int w = kx[k];
                h0 += px[0] * w; h1 += px[1] * w; h2 += px[2] * w;
            }
            const int w = ky[j];
            a0 += clip8(h0 >> AUG_PREC) * w; a1 += clip8(h1 >> AUG_PREC) * w; a2 += clip8(h2 >> AUG_PREC) * w;
        }
        int r = clip8(a0 >> AUG_PREC), g = clip8(a1 >> AUG_PREC), bl = clip8(a2 >> AUG_PREC);
        if (gray >= 0) { const int v = gray == 0 ? r : (gray == 1 ? g : bl); r = g = bl = v; }     // RandomGray
        pr[i] = (uint8_t)r; pg[i] = (uint8_t)g; pb[i] = (uint8_t)bl;
    }
    // 2. ColorJitter, per-frame order; each thread keeps to its own pixels, only the contrast mean crosses threads
    for (int o = 0; o < 4; ++o) {
        const int op = fp[1 + o];
        if (op < 0) break;
        const float alpha = __int_as_float(fp[5 + o]);
        const bool interp = alpha >= 0.f && alpha <= 1.f;
        if (op == 1) {                                   // contrast: degenerate = mean luma of the frame
            unsigned long long sum = 0;
            for (int i = threadIdx.x; i < npix; i += AUG_THREADS) sum += (unsigned)luma(pr[i], pg[i], pb[i]);
            for (int d = 16; d; d >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, d);
            if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = sum;
            __syncthreads();
            if (threadIdx.x == 0) {
                unsigned long long t = 0;
                for (int w = 0; w < AUG_THREADS / 32; ++w) t += red[w];
                s_mean = (int)__dadd_rn(__ddiv_rn((double)t, (double)npix), 0.5);       // int(ImageStat mean + 0.5)
            }
            __syncthreads();
        }
        const int mean = s_mean;
        const int shift = fp[9];
        for (int i = threadIdx.x; i < npix; i += AUG_THREADS) {
            int r = pr[i], g = pg[i], bl = pb[i];
            if (op == 0) {                               // brightness: blend(black, img)
                r = blend8(0, r, alpha, interp); g = blend8(0, g, alpha, interp); bl = blend8(0, bl, alpha, interp);
            } else if (op == 1) {
                r = blend8(mean, r, alpha, interp); g = blend8(mean, g, alpha, interp); bl = blend8(mean, bl, alpha, interp);
            } else if (op == 2) {                        // saturation: blend(luma, img)
                const int L = luma(r, g, bl);
                r = blend8(L, r, alpha, interp); g = blend8(L, g, alpha, interp); bl = blend8(L, bl, alpha, interp);
            } else {
                hue_rotate(r, g, bl, shift, lut_i, lut_f, lut_fs);
            }
            pr[i] = (uint8_t)r; pg[i] = (uint8_t)g; pb[i] = (uint8_t)bl;
        }
        __syncthreads();                                 // s_mean / red are reused by a later contrast step
    }
    // 3. ToTensor (/255) + Normalize ((x - mean) / std) through the per-byte table; block[b, n, c, t, :, :]
    __syncthreads();                                     // the vector loads below cross the per-thread pixel ownership
    const int n = f / p.SL, t = f - n * p.SL;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        float* dst = p.out + ((((size_t)b * p.N + n) * 3 + c) * p.SL + t) * npix;
        const uint8_t* pl = img + c * npix;
        const float* lut = lut_out[c];
        if ((npix & 3) == 0) {
            for (int i = threadIdx.x; i < (npix >> 2); i += AUG_THREADS) {
                const uchar4 v = reinterpret_cast<const uchar4*>(pl)[i];
                reinterpret_cast<float4*>(dst)[i] = make_float4(lut[v.x], lut[v.y], lut[v.z], lut[v.w]);
            }
        } else {
            for (int i = threadIdx.x; i < npix; i += AUG_THREADS) dst[i] = lut[pl[i]];
        }
    }
}

}  // namespace

extern "C" int dpc_augment_clips(const uint8_t* frames, const int32_t* tables, const int32_t* frame_params, const float* mean,
                                 const float* stdv, float* out, int B, int F, int H, int W, int Ho, int Wo, int K, int num_seq,
                                 int seq_len, void* stream) {
    DPC_REQUIRE(frames && tables && frame_params && mean && stdv && out, "dpc_augment_clips: null pointer");
    DPC_REQUIRE(B > 0 && F > 0 && H > 0 && W > 0 && Ho > 0 && Wo > 0 && K > 0, "dpc_augment_clips: bad dims");
    DPC_REQUIRE(num_seq * seq_len == F, "dpc_augment_clips: num_seq (%d) * seq_len (%d) must equal the frame count %d", num_seq,
                seq_len, F);
    const size_t smem = (size_t)3 * Ho * Wo;
    DPC_REQUIRE(smem <= 200 * 1024, "dpc_augment_clips: output frame %dx%d does not fit shared memory", Wo, Ho);
    AugParams p;
    p.frames = frames; p.tables = tables; p.fparams = frame_params; p.out = out;
    p.B = B; p.F = F; p.H = H; p.W = W; p.Ho = Ho; p.Wo = Wo; p.K = K; p.N = num_seq; p.SL = seq_len;
    for (int c = 0; c < 3; ++c) { p.mean[c] = mean[c]; p.stdv[c] = stdv[c]; }
    DPC_CUDA(cudaFuncSetAttribute(augment_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    augment_kernel<<<B * F, AUG_THREADS, smem, reinterpret_cast<cudaStream_t>(stream)>>>(p);
    DPC_LAUNCH_CHECK();
    return DPC_OK;
}

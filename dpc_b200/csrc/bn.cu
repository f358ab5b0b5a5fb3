// BatchNorm3d with batch statistics (track_running_stats=False), fused with ReLU / residual /
// max-pool / temporal average.  All kernels are single-pass HBM streams over channels-last rows:
// a thread owns one channel quad (float4) and walks rows, so per-channel constants stay in registers
// and every access is a coalesced 16-byte vector.  Per-channel reductions: fp32 over <=64 rows,
// then fp64 per thread / block, then fp64 atomics (one per channel per block).
// Replaces nn.BatchNorm3d + relu_ + `out += residual` (backbone/resnet_2d3d.py:55-78,91-114,
// 212-214,243) and F.avg_pool3d + relu (dpc/model_3d.py:53-57).
#include "common.cuh"
#include <cuda_bf16.h>

namespace {

constexpr int THREADS = 256;
constexpr int STRIP = 64;   // rows accumulated in fp32 before spilling into fp64

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }

// block-level: add this thread's 8 doubles (4 channels x {a,b}) into ws[2C] with one atomic per
// channel per block.
__device__ __forceinline__ void block_reduce_atomic(double a[4], double b[4], int cq, int rl, int C4,
                                                    int rg, double* ws, int C) {
    extern __shared__ double red[];   // [rg][C4*8]
    double* mine = red + ((size_t)rl * C4 + cq) * 8;
#pragma unroll
    for (int i = 0; i < 4; ++i) { mine[i] = a[i]; mine[4 + i] = b[i]; }
    __syncthreads();
    for (int idx = threadIdx.x; idx < C4 * 8; idx += blockDim.x) {
        double s = 0.0;
        for (int r = 0; r < rg; ++r) s += red[(size_t)r * C4 * 8 + idx];
        int q = idx / 8, e = idx % 8;
        int ch = q * 4 + (e & 3);
        atomicAdd(ws + (e < 4 ? 0 : C) + ch, s);
    }
}

__global__ void __launch_bounds__(THREADS) bn_stats_kernel(const float* __restrict__ y, long long rows,
                                                            int C, double* __restrict__ ws) {
    const int C4 = C / 4, rg = THREADS / C4;
    const int cq = threadIdx.x % C4, rl = threadIdx.x / C4;
    double s[4] = {0, 0, 0, 0}, ss[4] = {0, 0, 0, 0};
    const long long chunk = (long long)STRIP * rg;
    for (long long base = (long long)blockIdx.x * chunk; base < rows; base += (long long)gridDim.x * chunk) {
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
        long long end = base + chunk < rows ? base + chunk : rows;
        for (long long r = base + rl; r < end; r += rg) {
            float4 v = ld4(y + r * C + cq * 4);
            a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
            b.x = fmaf(v.x, v.x, b.x); b.y = fmaf(v.y, v.y, b.y); b.z = fmaf(v.z, v.z, b.z); b.w = fmaf(v.w, v.w, b.w);
        }
        s[0] += a.x; s[1] += a.y; s[2] += a.z; s[3] += a.w;
        ss[0] += b.x; ss[1] += b.y; ss[2] += b.z; ss[3] += b.w;
    }
    block_reduce_atomic(s, ss, cq, rl, C4, rg, ws, C);
}

__global__ void bn_finalize_kernel(const double* __restrict__ ws, long long rows, int C, float eps,
                                   float* __restrict__ mean, float* __restrict__ rstd) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    double m = ws[c] / (double)rows;
    double var = ws[C + c] / (double)rows - m * m;   // biased variance
    if (var < 0.0) var = 0.0;
    mean[c] = (float)m;
    rstd[c] = (float)(1.0 / sqrt(var + (double)eps));
}

struct Affine4 { float4 m, sc, b; };
__device__ __forceinline__ Affine4 make_affine(const float* mean, const float* rstd, const float* gamma,
                                               const float* beta, int c) {
    float4 r = ld4(rstd + c), g = ld4(gamma + c);
    Affine4 a;
    a.m = ld4(mean + c);
    a.b = ld4(beta + c);
    a.sc = make_float4(g.x * r.x, g.y * r.y, g.z * r.z, g.w * r.w);
    return a;
}
// (v - mean) * (gamma * rstd) + beta: subtract first, so a large mean does not cancel in fp32
__device__ __forceinline__ float4 affine(float4 v, const Affine4& a) {
    return make_float4(fmaf(v.x - a.m.x, a.sc.x, a.b.x), fmaf(v.y - a.m.y, a.sc.y, a.b.y),
                       fmaf(v.z - a.m.z, a.sc.z, a.b.z), fmaf(v.w - a.m.w, a.sc.w, a.b.w));
}

// ---- split-bf16 planes (tensor-core operand format): value ~= hi + lo ----------------------------
__device__ __forceinline__ float bf16_bits_to_float(uint32_t b) { return __uint_as_float(b << 16); }
__device__ __forceinline__ float4 unpack_bf16x4(uint2 p) {
    return make_float4(bf16_bits_to_float(p.x & 0xFFFFu), bf16_bits_to_float(p.x >> 16),
                       bf16_bits_to_float(p.y & 0xFFFFu), bf16_bits_to_float(p.y >> 16));
}
__device__ __forceinline__ float4 ld4_hi(const void* hi, long long off) {
    return unpack_bf16x4(*reinterpret_cast<const uint2*>(reinterpret_cast<const uint16_t*>(hi) + off));
}
__device__ __forceinline__ float4 ld4_planes(const void* hi, const void* lo, long long off) {
    float4 a = ld4_hi(hi, off), b = ld4_hi(lo, off);
    return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
}
__device__ __forceinline__ uint32_t f2bf_rn(float f) {      // round-to-nearest-even bf16 bits (finite inputs)
    uint32_t u = __float_as_uint(f);
    return (u + 0x7FFFu + ((u >> 16) & 1u)) >> 16;
}
// hi = bf16(v), lo = bf16(v - hi), packed pairwise with the hardware cvt.rn.bf16x2.f32 (round-to-nearest-even, the
// same rounding as dpc_split_bf16); ~14 instructions per float4 instead of ~50 with integer bit arithmetic
__device__ __forceinline__ void st4_planes(void* hi, void* lo, long long off, float4 v) {
    const __nv_bfloat162 h01 = __floats2bfloat162_rn(v.x, v.y), h23 = __floats2bfloat162_rn(v.z, v.w);
    const float2 f01 = __bfloat1622float2(h01), f23 = __bfloat1622float2(h23);
    const __nv_bfloat162 l01 = __floats2bfloat162_rn(v.x - f01.x, v.y - f01.y);
    const __nv_bfloat162 l23 = __floats2bfloat162_rn(v.z - f23.x, v.w - f23.y);
    *reinterpret_cast<uint2*>(reinterpret_cast<uint16_t*>(hi) + off) =
        make_uint2(*reinterpret_cast<const uint32_t*>(&h01), *reinterpret_cast<const uint32_t*>(&h23));
    *reinterpret_cast<uint2*>(reinterpret_cast<uint16_t*>(lo) + off) =
        make_uint2(*reinterpret_cast<const uint32_t*>(&l01), *reinterpret_cast<const uint32_t*>(&l23));
}

// RES: 0 none, 1 raw residual (fp32 rows or bf16 planes), 2 batch-normalised residual (downsample branch)
// Output: fp32 rows (`out`, nullable) and/or split-bf16 planes (`out_hi/out_lo`, nullable).
template <int RES, bool RELU>
__global__ void __launch_bounds__(THREADS) bn_apply_kernel(const float* __restrict__ y, const float* mean,
                                                            const float* rstd, const float* gamma,
                                                            const float* beta, const float* __restrict__ res,
                                                            const void* __restrict__ res_hi,
                                                            const void* __restrict__ res_lo,
                                                            const float* r_mean, const float* r_rstd,
                                                            const float* r_gamma, const float* r_beta,
                                                            float* __restrict__ out, void* __restrict__ out_hi,
                                                            void* __restrict__ out_lo, long long rows, int C) {
    const int C4 = C / 4, rg = THREADS / C4;
    const int cq = threadIdx.x % C4, rl = threadIdx.x / C4;
    const Affine4 a = make_affine(mean, rstd, gamma, beta, cq * 4);
    Affine4 ar;
    if (RES == 2) ar = make_affine(r_mean, r_rstd, r_gamma, r_beta, cq * 4);
    for (long long r = (long long)blockIdx.x * rg + rl; r < rows; r += (long long)gridDim.x * rg) {
        const long long off = r * C + cq * 4;
        float4 v = affine(ld4(y + off), a);
        if (RES == 1) {
            float4 q = res ? ld4(res + off) : ld4_planes(res_hi, res_lo, off);
            v.x += q.x; v.y += q.y; v.z += q.z; v.w += q.w;
        } else if (RES == 2) {
            float4 q = affine(ld4(res + off), ar);
            v.x += q.x; v.y += q.y; v.z += q.z; v.w += q.w;
        }
        if (RELU) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
        if (out) st4(out + off, v);
        if (out_hi) st4_planes(out_hi, out_lo, off, v);
    }
}

// ---- backward -------------------------------------------------------------------------------
// ReLU mask from the forward output: fp32 rows (`out`) or the hi plane (sign(hi) == sign(value))
__device__ __forceinline__ float4 masked_grad(const float* dout, const float* out, const void* out_hi,
                                              long long off, bool relu) {
    float4 g = ld4(dout + off);
    if (relu) {
        float4 o = out ? ld4(out + off) : ld4_hi(out_hi, off);
        g.x = o.x > 0.f ? g.x : 0.f; g.y = o.y > 0.f ? g.y : 0.f;
        g.z = o.z > 0.f ? g.z : 0.f; g.w = o.w > 0.f ? g.w : 0.f;
    }
    return g;
}

__global__ void __launch_bounds__(THREADS) bn_bwd_reduce_kernel(const float* __restrict__ dout,
                                                                 const float* __restrict__ out,
                                                                 const void* __restrict__ out_hi, int relu,
                                                                 const float* __restrict__ y, const float* mean,
                                                                 const float* rstd, long long rows, int C,
                                                                 double* __restrict__ ws) {
    const int C4 = C / 4, rg = THREADS / C4;
    const int cq = threadIdx.x % C4, rl = threadIdx.x / C4;
    const float4 m = ld4(mean + cq * 4), rs = ld4(rstd + cq * 4);
    double s[4] = {0, 0, 0, 0}, sx[4] = {0, 0, 0, 0};   // sum g, sum g*xhat
    const long long chunk = (long long)STRIP * rg;
    for (long long base = (long long)blockIdx.x * chunk; base < rows; base += (long long)gridDim.x * chunk) {
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
        long long end = base + chunk < rows ? base + chunk : rows;
        for (long long r = base + rl; r < end; r += rg) {
            const long long off = r * C + cq * 4;
            float4 g = masked_grad(dout, out, out_hi, off, relu != 0);
            float4 v = ld4(y + off);
            a.x += g.x; a.y += g.y; a.z += g.z; a.w += g.w;
            b.x = fmaf(g.x, (v.x - m.x) * rs.x, b.x); b.y = fmaf(g.y, (v.y - m.y) * rs.y, b.y);
            b.z = fmaf(g.z, (v.z - m.z) * rs.z, b.z); b.w = fmaf(g.w, (v.w - m.w) * rs.w, b.w);
        }
        s[0] += a.x; s[1] += a.y; s[2] += a.z; s[3] += a.w;
        sx[0] += b.x; sx[1] += b.y; sx[2] += b.z; sx[3] += b.w;
    }
    block_reduce_atomic(s, sx, cq, rl, C4, rg, ws, C);
}

__global__ void bn_bwd_finalize_kernel(const double* __restrict__ ws, int C, float* __restrict__ dgamma,
                                       float* __restrict__ dbeta) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    dbeta[c] = (float)ws[c];
    dgamma[c] = (float)ws[C + c];
}

__global__ void __launch_bounds__(THREADS) bn_bwd_apply_kernel(const float* __restrict__ dout,
                                                                const float* __restrict__ out,
                                                                const void* __restrict__ out_hi, int relu,
                                                                const float* __restrict__ y, const float* mean,
                                                                const float* rstd, const float* gamma,
                                                                const double* __restrict__ ws,
                                                                float* __restrict__ dy, void* __restrict__ dy_hi,
                                                                void* __restrict__ dy_lo, float* __restrict__ g_out,
                                                                long long rows, int C) {
    const int C4 = C / 4, rg = THREADS / C4;
    const int cq = threadIdx.x % C4, rl = threadIdx.x / C4;
    const int c = cq * 4;
    const float4 m = ld4(mean + c), rs = ld4(rstd + c), ga = ld4(gamma + c);
    const double inv_n = 1.0 / (double)rows;
    float mb[4], mg[4];   // dbeta/n, dgamma/n
#pragma unroll
    for (int i = 0; i < 4; ++i) { mb[i] = (float)(ws[c + i] * inv_n); mg[i] = (float)(ws[C + c + i] * inv_n); }
    const float k0 = ga.x * rs.x, k1 = ga.y * rs.y, k2 = ga.z * rs.z, k3 = ga.w * rs.w;
    for (long long r = (long long)blockIdx.x * rg + rl; r < rows; r += (long long)gridDim.x * rg) {
        const long long off = r * C + c;
        float4 g = masked_grad(dout, out, out_hi, off, relu != 0);
        float4 v = ld4(y + off);
        float4 d;
        d.x = k0 * (g.x - mb[0] - (v.x - m.x) * rs.x * mg[0]);
        d.y = k1 * (g.y - mb[1] - (v.y - m.y) * rs.y * mg[1]);
        d.z = k2 * (g.z - mb[2] - (v.z - m.z) * rs.z * mg[2]);
        d.w = k3 * (g.w - mb[3] - (v.w - m.w) * rs.w * mg[3]);
        if (dy) st4(dy + off, d);
        if (dy_hi) st4_planes(dy_hi, dy_lo, off, d);
        if (g_out) st4(g_out + off, g);
    }
}

// ---- stem tail: BN + ReLU + MaxPool (1,3,3)/(1,2,2)/(0,1,1) ------------------------------------
// Each thread produces TWO horizontally adjacent pooled outputs (wo0 = 2*wq, wo0 + 1): their windows share one of
// three input columns, so 15 loads / affine transforms serve two outputs instead of 18.
__global__ void __launch_bounds__(THREADS) bn_relu_maxpool_fwd_kernel(const float* __restrict__ y,
                                                                       const float* mean, const float* rstd,
                                                                       const float* gamma, const float* beta,
                                                                       float* __restrict__ out, int NT, int H,
                                                                       int W, int Ho, int Wo, int C) {
    const int C4 = C / 4, rg = THREADS / C4;
    const int cq = threadIdx.x % C4, rl = threadIdx.x / C4;
    const Affine4 a = make_affine(mean, rstd, gamma, beta, cq * 4);
    const int Wo2 = (Wo + 1) / 2;
    const long long rows = (long long)NT * Ho * Wo2;
    for (long long r = (long long)blockIdx.x * rg + rl; r < rows; r += (long long)gridDim.x * rg) {
        const int wq = (int)(r % Wo2);
        const int ho = (int)((r / Wo2) % Ho);
        const long long nt = r / ((long long)Wo2 * Ho);
        float4 b0 = make_float4(0.f, 0.f, 0.f, 0.f), b1 = b0;   // ReLU output is >= 0, so 0 is the identity
#pragma unroll
        for (int dh = -1; dh <= 1; ++dh) {
            const int h = 2 * ho + dh;
            if (h < 0 || h >= H) continue;
#pragma unroll
            for (int c = 0; c < 5; ++c) {
                const int w = 4 * wq - 1 + c;
                if (w < 0 || w >= W) continue;
                const float4 v = affine(ld4(y + ((nt * H + h) * W + w) * C + cq * 4), a);
                if (c <= 2) { b0.x = fmaxf(b0.x, v.x); b0.y = fmaxf(b0.y, v.y); b0.z = fmaxf(b0.z, v.z); b0.w = fmaxf(b0.w, v.w); }
                if (c >= 2) { b1.x = fmaxf(b1.x, v.x); b1.y = fmaxf(b1.y, v.y); b1.z = fmaxf(b1.z, v.z); b1.w = fmaxf(b1.w, v.w); }
            }
        }
        float* o = out + ((nt * Ho + ho) * Wo + 2 * wq) * C + cq * 4;
        st4(o, b0);
        if (2 * wq + 1 < Wo) st4(o + C, b1);
    }
}

// g[pos] = sum over pooling windows containing pos of dout[win] * [a(pos) == out[win] && a(pos) > 0]
__global__ void __launch_bounds__(THREADS) bn_relu_maxpool_bwd_kernel(const float* __restrict__ y,
                                                                       const float* mean, const float* rstd,
                                                                       const float* gamma, const float* beta,
                                                                       const float* __restrict__ out,
                                                                       const float* __restrict__ dout,
                                                                       float* __restrict__ g, int NT, int H, int W,
                                                                       int Ho, int Wo, int C) {
    const int C4 = C / 4, rg = THREADS / C4;
    const int cq = threadIdx.x % C4, rl = threadIdx.x / C4;
    const Affine4 a = make_affine(mean, rstd, gamma, beta, cq * 4);
    const long long rows = (long long)NT * H * W;
    for (long long r = (long long)blockIdx.x * rg + rl; r < rows; r += (long long)gridDim.x * rg) {
        int w = (int)(r % W);
        int h = (int)((r / W) % H);
        long long nt = r / ((long long)W * H);
        float4 v = affine(ld4(y + r * C + cq * 4), a);
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        // windows: ho in {h/2} for even h, {h/2, h/2+1} for odd h (same for w)
        int ho_lo = h >> 1, ho_hi = (h & 1) ? ho_lo + 1 : ho_lo;
        int wo_lo = w >> 1, wo_hi = (w & 1) ? wo_lo + 1 : wo_lo;
        for (int ho = ho_lo; ho <= ho_hi; ++ho) {
            if (ho >= Ho) continue;
            for (int wo = wo_lo; wo <= wo_hi; ++wo) {
                if (wo >= Wo) continue;
                long long o = ((nt * Ho + ho) * Wo + wo) * C + cq * 4;
                float4 p = ld4(out + o), d = ld4(dout + o);
                if (v.x > 0.f && v.x == p.x) acc.x += d.x;
                if (v.y > 0.f && v.y == p.y) acc.y += d.y;
                if (v.z > 0.f && v.z == p.z) acc.z += d.z;
                if (v.w > 0.f && v.w == p.w) acc.w += d.w;
            }
        }
        st4(g + r * C + cq * 4, acc);
    }
}

// ---- fused stem-tail backward: max-pool bwd + ReLU bwd + BN bwd without materialising g ----------
// g(row, c) is recomputed on the fly in both BN-backward passes (see bn_relu_maxpool_bwd_kernel).
__device__ __forceinline__ float4 stem_pool_grad(float4 v /* relu(bn(y)) pre-activation value */,
                                                 const float* __restrict__ out, const float* __restrict__ dout,
                                                 long long nt, int h, int w, int Ho, int Wo, int C, int cq) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    const int ho_lo = h >> 1, ho_hi = (h & 1) ? ho_lo + 1 : ho_lo;
    const int wo_lo = w >> 1, wo_hi = (w & 1) ? wo_lo + 1 : wo_lo;
    for (int ho = ho_lo; ho <= ho_hi; ++ho) {
        if (ho >= Ho) continue;
        for (int wo = wo_lo; wo <= wo_hi; ++wo) {
            if (wo >= Wo) continue;
            const long long o = ((nt * Ho + ho) * Wo + wo) * C + cq * 4;
            const float4 p = ld4(out + o), d = ld4(dout + o);
            if (v.x > 0.f && v.x == p.x) acc.x += d.x;
            if (v.y > 0.f && v.y == p.y) acc.y += d.y;
            if (v.z > 0.f && v.z == p.z) acc.z += d.z;
            if (v.w > 0.f && v.w == p.w) acc.w += d.w;
        }
    }
    return acc;
}

// One thread owns a 2x2 block of conv1-grid positions (x one channel quad): the four positions share the
// four pooling windows (ho in {h2, h2+1}) x (wo in {w2, w2+1}), so out/dout are loaded once per block
// (12 vector loads per 4 positions instead of 22).
struct TailBlock {
    float4 yv[4];      // y at (2h2+a, 2w2+b), index a*2+b
    float4 g[4];       // gradient w.r.t. bn(y) at those positions
    bool ok[4];
};
__device__ __forceinline__ void tail_block(TailBlock& tb, const float* __restrict__ y, const Affine4& a,
                                           const float* __restrict__ out, const float* __restrict__ dout, long long nt,
                                           int h2, int w2, int H, int W, int Ho, int Wo, int C, int cq) {
    float4 p[4], d[4];
    bool wok[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int ho = h2 + (i >> 1), wo = w2 + (i & 1);
        wok[i] = ho < Ho && wo < Wo;
        if (wok[i]) {
            const long long o = ((nt * Ho + ho) * Wo + wo) * C + cq * 4;
            p[i] = ld4(out + o); d[i] = ld4(dout + o);
        } else {
            p[i] = make_float4(-1.f, -1.f, -1.f, -1.f); d[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int aa = i >> 1, bb = i & 1;
        const int h = 2 * h2 + aa, w = 2 * w2 + bb;
        tb.ok[i] = h < H && w < W;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        if (tb.ok[i]) {
            tb.yv[i] = ld4(y + ((nt * H + h) * W + w) * C + cq * 4);
            const float4 v = affine(tb.yv[i], a);
            // windows containing (h, w): ho = h2 (+1 if the row is odd), wo = w2 (+1 if the column is odd)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int dh = j >> 1, dw = j & 1;
                if ((dh && !aa) || (dw && !bb)) continue;
                if (v.x > 0.f && v.x == p[j].x) acc.x += d[j].x;
                if (v.y > 0.f && v.y == p[j].y) acc.y += d[j].y;
                if (v.z > 0.f && v.z == p[j].z) acc.z += d[j].z;
                if (v.w > 0.f && v.w == p[j].w) acc.w += d[j].w;
            }
        } else {
            tb.yv[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        tb.g[i] = acc;
    }
}

// The reduce pass on the POOLED grid (4x fewer elements, y is not read at all): the gradient of a pooling
// window lands on its arg-max position, whose normalised value follows from the pooled output itself,
//   a = gamma * xhat + beta  ->  xhat = (a - beta) / gamma            (a > 0: ReLU passed)
// so  sum g = sum_w dout_w [a_w > 0]  and  sum g*xhat = sum_w dout_w [a_w > 0] (a_w - beta) / gamma.
// (Requires gamma != 0; exact ties inside a window have measure zero.)
__global__ void __launch_bounds__(THREADS) stem_tail_bwd_reduce_pooled_kernel(
    const float* gamma, const float* beta, const float* __restrict__ out, const float* __restrict__ dout,
    long long rows, int C, double* __restrict__ ws) {
    const int C4 = C / 4, rg = THREADS / C4;
    const int cq = threadIdx.x % C4, rl = threadIdx.x / C4;
    const float4 ga = ld4(gamma + cq * 4), be = ld4(beta + cq * 4);
    const float4 ig = make_float4(1.f / ga.x, 1.f / ga.y, 1.f / ga.z, 1.f / ga.w);
    double s[4] = {0, 0, 0, 0}, sx[4] = {0, 0, 0, 0};
    const long long chunk = (long long)STRIP * rg;
    for (long long base = (long long)blockIdx.x * chunk; base < rows; base += (long long)gridDim.x * chunk) {
        float4 pa = make_float4(0.f, 0.f, 0.f, 0.f), pb = pa;
        long long end = base + chunk < rows ? base + chunk : rows;
        for (long long r = base + rl; r < end; r += rg) {
            const long long off = r * C + cq * 4;
            const float4 a = ld4(out + off), d = ld4(dout + off);
            const float gx = a.x > 0.f ? d.x : 0.f, gy = a.y > 0.f ? d.y : 0.f;
            const float gz = a.z > 0.f ? d.z : 0.f, gw = a.w > 0.f ? d.w : 0.f;
            pa.x += gx; pa.y += gy; pa.z += gz; pa.w += gw;
            pb.x = fmaf(gx, (a.x - be.x) * ig.x, pb.x); pb.y = fmaf(gy, (a.y - be.y) * ig.y, pb.y);
            pb.z = fmaf(gz, (a.z - be.z) * ig.z, pb.z); pb.w = fmaf(gw, (a.w - be.w) * ig.w, pb.w);
        }
        s[0] += pa.x; s[1] += pa.y; s[2] += pa.z; s[3] += pa.w;
        sx[0] += pb.x; sx[1] += pb.y; sx[2] += pb.z; sx[3] += pb.w;
    }
    block_reduce_atomic(s, sx, cq, rl, C4, rg, ws, C);
}

__global__ void __launch_bounds__(THREADS, 3) stem_tail_bwd_reduce_kernel(
    const float* __restrict__ y, const float* mean, const float* rstd, const float* gamma, const float* beta,
    const float* __restrict__ out, const float* __restrict__ dout, int NT, int H, int W, int Ho, int Wo, int C,
    double* __restrict__ ws) {
    const int C4 = C / 4, rg = THREADS / C4;
    const int cq = threadIdx.x % C4, rl = threadIdx.x / C4;
    const Affine4 a = make_affine(mean, rstd, gamma, beta, cq * 4);
    const float4 m = ld4(mean + cq * 4), rs = ld4(rstd + cq * 4);
    const int H2 = (H + 1) / 2, W2 = (W + 1) / 2;
    const long long blocks = (long long)NT * H2 * W2;
    double s[4] = {0, 0, 0, 0}, sx[4] = {0, 0, 0, 0};
    const long long chunk = (long long)(STRIP / 4) * rg;           // 16 blocks = 64 positions per fp32 strip
    for (long long base = (long long)blockIdx.x * chunk; base < blocks; base += (long long)gridDim.x * chunk) {
        float4 pa = make_float4(0.f, 0.f, 0.f, 0.f), pb = pa;
        long long end = base + chunk < blocks ? base + chunk : blocks;
        for (long long r = base + rl; r < end; r += rg) {
            const int w2 = (int)(r % W2), h2 = (int)((r / W2) % H2);
            const long long nt = r / ((long long)W2 * H2);
            TailBlock tb;
            tail_block(tb, y, a, out, dout, nt, h2, w2, H, W, Ho, Wo, C, cq);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float4 g = tb.g[i], yv = tb.yv[i];          // g == 0 where the position is out of range
                pa.x += g.x; pa.y += g.y; pa.z += g.z; pa.w += g.w;
                pb.x = fmaf(g.x, (yv.x - m.x) * rs.x, pb.x); pb.y = fmaf(g.y, (yv.y - m.y) * rs.y, pb.y);
                pb.z = fmaf(g.z, (yv.z - m.z) * rs.z, pb.z); pb.w = fmaf(g.w, (yv.w - m.w) * rs.w, pb.w);
            }
        }
        s[0] += pa.x; s[1] += pa.y; s[2] += pa.z; s[3] += pa.w;
        sx[0] += pb.x; sx[1] += pb.y; sx[2] += pb.z; sx[3] += pb.w;
    }
    block_reduce_atomic(s, sx, cq, rl, C4, rg, ws, C);
}

__global__ void __launch_bounds__(THREADS, 3) stem_tail_bwd_apply_kernel(
    const float* __restrict__ y, const float* mean, const float* rstd, const float* gamma, const float* beta,
    const float* __restrict__ out, const float* __restrict__ dout, int NT, int H, int W, int Ho, int Wo, int C,
    const double* __restrict__ ws, float* __restrict__ dy, void* __restrict__ dy_hi, void* __restrict__ dy_lo) {
    const int C4 = C / 4, rg = THREADS / C4;
    const int cq = threadIdx.x % C4, rl = threadIdx.x / C4;
    const int c = cq * 4;
    const Affine4 a = make_affine(mean, rstd, gamma, beta, c);
    const float4 m = ld4(mean + c), rs = ld4(rstd + c), ga = ld4(gamma + c);
    const double inv_n = 1.0 / ((double)NT * H * W);
    float mb[4], mg[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) { mb[i] = (float)(ws[c + i] * inv_n); mg[i] = (float)(ws[C + c + i] * inv_n); }
    const float k0 = ga.x * rs.x, k1 = ga.y * rs.y, k2 = ga.z * rs.z, k3 = ga.w * rs.w;
    const int H2 = (H + 1) / 2, W2 = (W + 1) / 2;
    const long long blocks = (long long)NT * H2 * W2;
    for (long long r = (long long)blockIdx.x * rg + rl; r < blocks; r += (long long)gridDim.x * rg) {
        const int w2 = (int)(r % W2), h2 = (int)((r / W2) % H2);
        const long long nt = r / ((long long)W2 * H2);
        TailBlock tb;
        tail_block(tb, y, a, out, dout, nt, h2, w2, H, W, Ho, Wo, C, cq);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            if (!tb.ok[i]) continue;
            const int h = 2 * h2 + (i >> 1), w = 2 * w2 + (i & 1);
            const long long off = ((nt * H + h) * W + w) * C + c;
            const float4 g = tb.g[i], yv = tb.yv[i];
            float4 d;
            d.x = k0 * (g.x - mb[0] - (yv.x - m.x) * rs.x * mg[0]);
            d.y = k1 * (g.y - mb[1] - (yv.y - m.y) * rs.y * mg[1]);
            d.z = k2 * (g.z - mb[2] - (yv.z - m.z) * rs.z * mg[2]);
            d.w = k3 * (g.w - mb[3] - (yv.w - m.w) * rs.w * mg[3]);
            if (dy) st4(dy + off, d);
            if (dy_hi) st4_planes(dy_hi, dy_lo, off, d);
        }
    }
}

// ---- temporal average + ReLU split ------------------------------------------------------------
__global__ void pool_split_fwd_kernel(const float* __restrict__ z, float* __restrict__ finf,
                                      float* __restrict__ feat, long long NB, int T, long long SC4) {
    // z [NB][T][S*C], outputs [NB][S*C]; SC4 = S*C/4
    long long total = NB * SC4;
    const float inv = 1.f / (float)T;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        long long n = i / SC4, e = i % SC4;
        float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int t = 0; t < T; ++t) {
            float4 v = ld4(z + ((n * T + t) * SC4 + e) * 4);
            s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
        }
        s.x *= inv; s.y *= inv; s.z *= inv; s.w *= inv;
        st4(finf + i * 4, s);
        st4(feat + i * 4, make_float4(fmaxf(s.x, 0.f), fmaxf(s.y, 0.f), fmaxf(s.z, 0.f), fmaxf(s.w, 0.f)));
    }
}

__global__ void pool_split_bwd_kernel(const float* __restrict__ finf, const float* __restrict__ dfinf,
                                      const float* __restrict__ dfeat, float* __restrict__ dz, long long NB,
                                      int T, long long SC4) {
    long long total = NB * SC4;
    const float inv = 1.f / (float)T;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        long long n = i / SC4, e = i % SC4;
        float4 f = ld4(finf + i * 4);
        float4 a = dfinf ? ld4(dfinf + i * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
        float4 b = dfeat ? ld4(dfeat + i * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
        float4 d;
        d.x = (a.x + (f.x > 0.f ? b.x : 0.f)) * inv;
        d.y = (a.y + (f.y > 0.f ? b.y : 0.f)) * inv;
        d.z = (a.z + (f.z > 0.f ? b.z : 0.f)) * inv;
        d.w = (a.w + (f.w > 0.f ? b.w : 0.f)) * inv;
        for (int t = 0; t < T; ++t) st4(dz + ((n * T + t) * SC4 + e) * 4, d);
    }
}

int check_c(int C, const char* who) {
    DPC_REQUIRE(C >= 4 && C % 4 == 0 && (THREADS % (C / 4)) == 0, "%s: unsupported channel count %d", who, C);
    return DPC_OK;
}
int stream_grid(long long rows, int rg) {
    long long need = (rows + rg - 1) / rg;
    long long cap = (long long)dpc_num_sms() * 16;
    return (int)(need < cap ? (need > 0 ? need : 1) : cap);
}

}  // namespace

extern "C" int dpc_bn_stats(const float* y, int64_t rows, int C, double* ws, float* mean, float* rstd,
                            float eps, void* stream) {
    DPC_REQUIRE(y && ws && mean && rstd && rows > 0, "dpc_bn_stats: bad args");
    if (int rc = check_c(C, "dpc_bn_stats")) return rc;
    cudaStream_t st = as_stream(stream);
    const int C4 = C / 4, rg = THREADS / C4;
    DPC_CUDA(cudaMemsetAsync(ws, 0, sizeof(double) * 2 * C, st));
    long long chunks = (rows + (long long)STRIP * rg - 1) / ((long long)STRIP * rg);
    int grid = (int)(chunks < (long long)dpc_num_sms() * 8 ? chunks : (long long)dpc_num_sms() * 8);
    size_t smem = sizeof(double) * (size_t)rg * C4 * 8;
    bn_stats_kernel<<<grid, THREADS, smem, st>>>(y, rows, C, ws);
    DPC_LAUNCH_CHECK();
    bn_finalize_kernel<<<ceil_div(C, 128), 128, 0, st>>>(ws, rows, C, eps, mean, rstd);
    DPC_LAUNCH_CHECK();
    return DPC_OK;
}

// mean / rstd from sums accumulated elsewhere (the conv epilogue): ws = [sum[C] | sumsq[C]] doubles
extern "C" int dpc_bn_finalize(const double* ws, int64_t rows, int C, float eps, float* mean, float* rstd, void* stream) {
    DPC_REQUIRE(ws && mean && rstd && rows > 0 && C > 0, "dpc_bn_finalize: bad args");
    bn_finalize_kernel<<<ceil_div(C, 128), 128, 0, as_stream(stream)>>>(ws, rows, C, eps, mean, rstd);
    DPC_LAUNCH_CHECK();
    return DPC_OK;
}

extern "C" int dpc_bn_apply_fwd(const float* y, const float* mean, const float* rstd, const float* gamma,
                                const float* beta, const float* res, const void* res_hi, const void* res_lo,
                                const float* r_mean, const float* r_rstd, const float* r_gamma,
                                const float* r_beta, int relu, float* out, void* out_hi, void* out_lo,
                                int64_t rows, int C, void* stream) {
    DPC_REQUIRE(y && mean && rstd && gamma && beta && rows > 0, "dpc_bn_apply_fwd: bad args");
    DPC_REQUIRE(out || (out_hi && out_lo), "dpc_bn_apply_fwd: no output");
    DPC_REQUIRE(!out_hi == !out_lo && !res_hi == !res_lo, "dpc_bn_apply_fwd: planes come in pairs");
    DPC_REQUIRE(!(res && res_hi), "dpc_bn_apply_fwd: residual given twice");
    if (int rc = check_c(C, "dpc_bn_apply_fwd")) return rc;
    const int mode = (res || res_hi) ? (r_mean ? 2 : 1) : 0;
    if (mode == 2) DPC_REQUIRE(res, "dpc_bn_apply_fwd: the batch-normalised residual must be fp32 rows");
    if (mode == 2) DPC_REQUIRE(r_rstd && r_gamma && r_beta, "dpc_bn_apply_fwd: incomplete residual BN");
    cudaStream_t st = as_stream(stream);
    const int rg = THREADS / (C / 4);
    const int grid = stream_grid(rows, rg);
#define LAUNCH(R, L) bn_apply_kernel<R, L><<<grid, THREADS, 0, st>>>(y, mean, rstd, gamma, beta, res, res_hi, res_lo, r_mean, r_rstd, r_gamma, r_beta, out, out_hi, out_lo, rows, C)
    if (mode == 0) { if (relu) LAUNCH(0, true); else LAUNCH(0, false); }
    else if (mode == 1) { if (relu) LAUNCH(1, true); else LAUNCH(1, false); }
    else { if (relu) LAUNCH(2, true); else LAUNCH(2, false); }
#undef LAUNCH
    DPC_LAUNCH_CHECK();
    return DPC_OK;
}

extern "C" int dpc_bn_bwd(const float* dout, const float* out, const void* out_hi, int relu, const float* y,
                          const float* mean, const float* rstd, const float* gamma, double* ws, float* dgamma,
                          float* dbeta, float* dy, void* dy_hi, void* dy_lo, float* g_out, int64_t rows, int C,
                          void* stream) {
    DPC_REQUIRE(dout && y && mean && rstd && gamma && ws && dgamma && dbeta && rows > 0, "dpc_bn_bwd: bad args");
    DPC_REQUIRE(dy || (dy_hi && dy_lo), "dpc_bn_bwd: no output");
    DPC_REQUIRE(!dy_hi == !dy_lo, "dpc_bn_bwd: planes come in pairs");
    DPC_REQUIRE(!relu || out || out_hi, "dpc_bn_bwd: relu needs the forward output (rows or hi plane)");
    if (int rc = check_c(C, "dpc_bn_bwd")) return rc;
    cudaStream_t st = as_stream(stream);
    const int C4 = C / 4, rg = THREADS / C4;
    DPC_CUDA(cudaMemsetAsync(ws, 0, sizeof(double) * 2 * C, st));
    long long chunks = (rows + (long long)STRIP * rg - 1) / ((long long)STRIP * rg);
    int grid = (int)(chunks < (long long)dpc_num_sms() * 8 ? chunks : (long long)dpc_num_sms() * 8);
    size_t smem = sizeof(double) * (size_t)rg * C4 * 8;
    bn_bwd_reduce_kernel<<<grid, THREADS, smem, st>>>(dout, out, out_hi, relu, y, mean, rstd, rows, C, ws);
    DPC_LAUNCH_CHECK();
    bn_bwd_finalize_kernel<<<ceil_div(C, 128), 128, 0, st>>>(ws, C, dgamma, dbeta);
    DPC_LAUNCH_CHECK();
    bn_bwd_apply_kernel<<<stream_grid(rows, rg), THREADS, 0, st>>>(dout, out, out_hi, relu, y, mean, rstd, gamma, ws,
                                                                   dy, dy_hi, dy_lo, g_out, rows, C);
    DPC_LAUNCH_CHECK();
    return DPC_OK;
}

// BatchNorm backward when the reduction already exists (ws = [sum g | sum g*xhat] doubles, e.g. produced by
// dpc_conv3d_dgrad_bnred_tc's epilogue): dgamma / dbeta from ws, then the apply pass only.
extern "C" int dpc_bn_bwd_apply(const float* dout, const float* out, const void* out_hi, int relu, const float* y,
                                const float* mean, const float* rstd, const float* gamma, const double* ws,
                                float* dgamma, float* dbeta, float* dy, void* dy_hi, void* dy_lo, float* g_out,
                                int64_t rows, int C, void* stream) {
    DPC_REQUIRE(dout && y && mean && rstd && gamma && ws && dgamma && dbeta && rows > 0, "dpc_bn_bwd_apply: bad args");
    DPC_REQUIRE(dy || (dy_hi && dy_lo), "dpc_bn_bwd_apply: no output");
    DPC_REQUIRE(!dy_hi == !dy_lo, "dpc_bn_bwd_apply: planes come in pairs");
    DPC_REQUIRE(!relu || out || out_hi, "dpc_bn_bwd_apply: relu needs the forward output (rows or hi plane)");
    if (int rc = check_c(C, "dpc_bn_bwd_apply")) return rc;
    cudaStream_t st = as_stream(stream);
    const int C4 = C / 4, rg = THREADS / C4;
    bn_bwd_finalize_kernel<<<ceil_div(C, 128), 128, 0, st>>>(ws, C, dgamma, dbeta);
    DPC_LAUNCH_CHECK();
    bn_bwd_apply_kernel<<<stream_grid(rows, rg), THREADS, 0, st>>>(dout, out, out_hi, relu, y, mean, rstd, gamma, ws,
                                                                   dy, dy_hi, dy_lo, g_out, rows, C);
    DPC_LAUNCH_CHECK();
    return DPC_OK;
}


extern "C" int dpc_bn_relu_maxpool_fwd(const float* y, const float* mean, const float* rstd, const float* gamma,
                                       const float* beta, float* out, int NT, int H, int W, int C, void* stream) {
    DPC_REQUIRE(y && mean && rstd && gamma && beta && out && NT > 0 && H > 0 && W > 0, "dpc_bn_relu_maxpool_fwd: bad args");
    if (int rc = check_c(C, "dpc_bn_relu_maxpool_fwd")) return rc;
    const int Ho = (H + 2 - 3) / 2 + 1, Wo = (W + 2 - 3) / 2 + 1;
    const int rg = THREADS / (C / 4);
    bn_relu_maxpool_fwd_kernel<<<stream_grid((long long)NT * Ho * ((Wo + 1) / 2), rg), THREADS, 0, as_stream(stream)>>>(
        y, mean, rstd, gamma, beta, out, NT, H, W, Ho, Wo, C);
    DPC_LAUNCH_CHECK();
    return DPC_OK;
}

extern "C" int dpc_bn_relu_maxpool_bwd(const float* y, const float* mean, const float* rstd, const float* gamma,
                                       const float* beta, const float* out, const float* dout, float* g,
                                       int NT, int H, int W, int C, void* stream) {
    DPC_REQUIRE(y && mean && rstd && gamma && beta && out && dout && g && NT > 0, "dpc_bn_relu_maxpool_bwd: bad args");
    if (int rc = check_c(C, "dpc_bn_relu_maxpool_bwd")) return rc;
    const int Ho = (H + 2 - 3) / 2 + 1, Wo = (W + 2 - 3) / 2 + 1;
    const int rg = THREADS / (C / 4);
    bn_relu_maxpool_bwd_kernel<<<stream_grid((long long)NT * H * W, rg), THREADS, 0, as_stream(stream)>>>(
        y, mean, rstd, gamma, beta, out, dout, g, NT, H, W, Ho, Wo, C);
    DPC_LAUNCH_CHECK();
    return DPC_OK;
}

extern "C" int dpc_stem_tail_bwd(const float* y, const float* mean, const float* rstd, const float* gamma,
                                 const float* beta, const float* out, const float* dout, double* ws, float* dgamma,
                                 float* dbeta, float* dy, void* dy_hi, void* dy_lo, int NT, int H, int W, int C,
                                 int pooled_reduce, void* stream) {
    DPC_REQUIRE(y && mean && rstd && gamma && beta && out && dout && ws && dgamma && dbeta && NT > 0,
                "dpc_stem_tail_bwd: bad args");
    DPC_REQUIRE(dy || (dy_hi && dy_lo), "dpc_stem_tail_bwd: no output");
    DPC_REQUIRE(!dy_hi == !dy_lo, "dpc_stem_tail_bwd: planes come in pairs");
    if (int rc = check_c(C, "dpc_stem_tail_bwd")) return rc;
    cudaStream_t st = as_stream(stream);
    const int Ho = (H + 2 - 3) / 2 + 1, Wo = (W + 2 - 3) / 2 + 1;
    const int C4 = C / 4, rg = THREADS / C4;
    const long long blocks = (long long)NT * ((H + 1) / 2) * ((W + 1) / 2);
    DPC_CUDA(cudaMemsetAsync(ws, 0, sizeof(double) * 2 * C, st));
    size_t smem = sizeof(double) * (size_t)rg * C4 * 8;
    if (pooled_reduce) {
        const long long prow = (long long)NT * Ho * Wo;
        long long chunks = (prow + (long long)STRIP * rg - 1) / ((long long)STRIP * rg);
        int grid = (int)(chunks < (long long)dpc_num_sms() * 8 ? chunks : (long long)dpc_num_sms() * 8);
        stem_tail_bwd_reduce_pooled_kernel<<<grid, THREADS, smem, st>>>(gamma, beta, out, dout, prow, C, ws);
    } else {
        const long long per = (long long)(STRIP / 4) * rg;
        long long chunks = (blocks + per - 1) / per;
        int grid = (int)(chunks < (long long)dpc_num_sms() * 8 ? chunks : (long long)dpc_num_sms() * 8);
        stem_tail_bwd_reduce_kernel<<<grid, THREADS, smem, st>>>(y, mean, rstd, gamma, beta, out, dout, NT, H, W, Ho, Wo, C, ws);
    }
    DPC_LAUNCH_CHECK();
    bn_bwd_finalize_kernel<<<ceil_div(C, 128), 128, 0, st>>>(ws, C, dgamma, dbeta);
    DPC_LAUNCH_CHECK();
    stem_tail_bwd_apply_kernel<<<stream_grid(blocks, rg), THREADS, 0, st>>>(y, mean, rstd, gamma, beta, out, dout, NT, H, W,
                                                                            Ho, Wo, C, ws, dy, dy_hi, dy_lo);
    DPC_LAUNCH_CHECK();
    return DPC_OK;
}

extern "C" int dpc_pool_split_fwd(const float* z, float* finf, float* feat, int NB, int T, int S, int C, void* stream) {
    DPC_REQUIRE(z && finf && feat && NB > 0 && T > 0 && S > 0 && C % 4 == 0, "dpc_pool_split_fwd: bad args");
    long long SC4 = (long long)S * C / 4, total = (long long)NB * SC4;
    int grid = (int)((total + 255) / 256);
    if (grid > dpc_num_sms() * 16) grid = dpc_num_sms() * 16;
    pool_split_fwd_kernel<<<grid, 256, 0, as_stream(stream)>>>(z, finf, feat, NB, T, SC4);
    DPC_LAUNCH_CHECK();
    return DPC_OK;
}

extern "C" int dpc_pool_split_bwd(const float* finf, const float* dfinf, const float* dfeat, float* dz,
                                  int NB, int T, int S, int C, void* stream) {
    DPC_REQUIRE(finf && dz && (dfinf || dfeat) && NB > 0 && T > 0 && S > 0 && C % 4 == 0, "dpc_pool_split_bwd: bad args");
    long long SC4 = (long long)S * C / 4, total = (long long)NB * SC4;
    int grid = (int)((total + 255) / 256);
    if (grid > dpc_num_sms() * 16) grid = dpc_num_sms() * 16;
    pool_split_bwd_kernel<<<grid, 256, 0, as_stream(stream)>>>(finf, dfinf, dfeat, dz, NB, T, SC4);
    DPC_LAUNCH_CHECK();
    return DPC_OK;
}

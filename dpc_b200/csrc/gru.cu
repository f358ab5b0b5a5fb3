// ConvGRU cell (kernel_size = 1) gate math, dropout, and small elementwise helpers.
// Replaces torch.sigmoid/tanh/mul/add and nn.Dropout at backbone/convrnn.py:29-33,78, and the
// ReLU of network_pred (dpc/model_3d.py:38,70).  With k = 1 every (b, l) row is an independent
// sequence, so these are plain coalesced row streams; the GEMMs around them are dpc_gemm_f32.
#include "common.cuh"

namespace {

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

// counter-based uniform in [0,1): splitmix64 of (seed, index)
__device__ __forceinline__ float uniform01(uint64_t seed, uint64_t idx) {
    uint64_t z = seed + 0x9E3779B97F4A7C15ull * (idx + 1);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z = z ^ (z >> 31);
    return (float)(z >> 40) * (1.0f / 16777216.0f);
}

__global__ void gru_gates_zr_kernel(const float* __restrict__ xz, const float* __restrict__ xr, int ldx,
                                    const float* __restrict__ hzr, const float* __restrict__ bz,
                                    const float* __restrict__ br, const float* __restrict__ h,
                                    float* __restrict__ z, float* __restrict__ r, float* __restrict__ hr,
                                    long long R, int D) {
    long long total = R * D;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        long long row = i / D;
        int c = (int)(i % D);
        float zz = sigmoidf_(xz[row * ldx + c] + hzr[row * 2 * D + c] + bz[c]);
        float rr = sigmoidf_(xr[row * ldx + c] + hzr[row * 2 * D + D + c] + br[c]);
        z[i] = zz;
        r[i] = rr;
        hr[i] = h[i] * rr;
    }
}

__global__ void gru_out_kernel(const float* __restrict__ xo, int ldx, const float* __restrict__ ho,
                               const float* __restrict__ bo, const float* __restrict__ h,
                               const float* __restrict__ z, float* __restrict__ o, float* __restrict__ hout,
                               float* __restrict__ keep, float p, uint64_t seed, uint64_t offset,
                               long long R, int D) {
    long long total = R * D;
    const float scale = p > 0.f ? 1.f / (1.f - p) : 1.f;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        long long row = i / D;
        int c = (int)(i % D);
        float oo = tanhf(xo[row * ldx + c] + ho[i] + bo[c]);
        float zz = z[i];
        float hn = h[i] * (1.f - zz) + oo * zz;
        float k = 1.f;
        if (p > 0.f) k = uniform01(seed, offset + (uint64_t)i) >= p ? scale : 0.f;
        o[i] = oo;
        hout[i] = hn * k;
        if (keep) keep[i] = k;
    }
}

__global__ void gru_bwd_out_kernel(const float* __restrict__ dhout, const float* __restrict__ keep,
                                   const float* __restrict__ h, const float* __restrict__ z,
                                   const float* __restrict__ o, float* __restrict__ dpre_o,
                                   float* __restrict__ dz_partial, float* __restrict__ dh, long long total) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        float d = dhout[i];
        if (keep) d *= keep[i];
        float zz = z[i], oo = o[i], hh = h[i];
        dpre_o[i] = d * zz * (1.f - oo * oo);
        dz_partial[i] = d * (oo - hh);
        dh[i] = d * (1.f - zz);
    }
}

__global__ void gru_bwd_zr_kernel(const float* __restrict__ dhr, const float* __restrict__ h,
                                  const float* __restrict__ r, const float* __restrict__ z,
                                  const float* __restrict__ dz_partial, float* __restrict__ dpre_zr,
                                  float* __restrict__ dh, long long R, int D) {
    long long total = R * D;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        long long row = i / D;
        int c = (int)(i % D);
        float g = dhr[i], rr = r[i], zz = z[i];
        dpre_zr[row * 2 * D + c] = dz_partial[i] * zz * (1.f - zz);
        dpre_zr[row * 2 * D + D + c] = g * h[i] * rr * (1.f - rr);
        dh[i] += g * rr;
    }
}

// x and y may alias (in-place)
__global__ void bias_relu_kernel(const float* x, const float* __restrict__ b, float* y, int relu,
                                 long long R, int D) {
    long long total = R * D;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        float v = x[i] + (b ? b[i % D] : 0.f);
        y[i] = relu ? fmaxf(v, 0.f) : v;
    }
}

// dy and dx may alias (in-place)
__global__ void relu_bwd_kernel(const float* __restrict__ y, const float* dy, float* dx, int accumulate,
                                long long n) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x) {
        float g = y[i] > 0.f ? dy[i] : 0.f;
        dx[i] = accumulate ? dx[i] + g : g;
    }
}

__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                            float* __restrict__ v, long long n, float lr, float b1, float b2, float eps,
                            float wd, float bc1, float bc2_sqrt, float gscale) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x) {
        float pi = p[i];
        float gi = g[i] * gscale + wd * pi;          // L2 weight decay folded into the gradient
        float mi = b1 * m[i] + (1.f - b1) * gi;
        float vi = b2 * v[i] + (1.f - b2) * gi * gi;
        m[i] = mi;
        v[i] = vi;
        float denom = sqrtf(vi) / bc2_sqrt + eps;
        p[i] = pi - (lr / bc1) * (mi / denom);
    }
}

int ew_grid(long long n) {
    long long b = (n + 255) / 256;
    long long cap = (long long)dpc_num_sms() * 16;
    return (int)(b < cap ? (b > 0 ? b : 1) : cap);
}

}  // namespace

extern "C" int dpc_gru_gates_zr(const float* xz, const float* xr, int ldx, const float* hzr, const float* bz,
                                const float* br, const float* h, float* z, float* r, float* hr, int64_t R,
                                int D, void* stream) {
    DPC_REQUIRE(xz && xr && hzr && bz && br && h && z && r && hr && R > 0 && D > 0, "dpc_gru_gates_zr: bad args");
    gru_gates_zr_kernel<<<ew_grid(R * D), 256, 0, as_stream(stream)>>>(xz, xr, ldx, hzr, bz, br, h, z, r, hr, R, D);
    DPC_LAUNCH_CHECK();
    return DPC_OK;
}

extern "C" int dpc_gru_out(const float* xo, int ldx, const float* ho, const float* bo, const float* h,
                           const float* z, float* o, float* hout, float* keep, float p, uint64_t seed,
                           uint64_t offset, int64_t R, int D, void* stream) {
    DPC_REQUIRE(xo && ho && bo && h && z && o && hout && R > 0 && D > 0, "dpc_gru_out: bad args");
    DPC_REQUIRE(p >= 0.f && p < 1.f, "dpc_gru_out: dropout p out of range");
    gru_out_kernel<<<ew_grid(R * D), 256, 0, as_stream(stream)>>>(xo, ldx, ho, bo, h, z, o, hout, keep, p, seed,
                                                                 offset, R, D);
    DPC_LAUNCH_CHECK();
    return DPC_OK;
}

extern "C" int dpc_gru_bwd_out(const float* dhout, const float* keep, const float* h, const float* z,
                               const float* o, float* dpre_o, float* dpre_z_partial, float* dh, int64_t R,
                               int D, void* stream) {
    DPC_REQUIRE(dhout && h && z && o && dpre_o && dpre_z_partial && dh && R > 0 && D > 0, "dpc_gru_bwd_out: bad args");
    gru_bwd_out_kernel<<<ew_grid(R * D), 256, 0, as_stream(stream)>>>(dhout, keep, h, z, o, dpre_o, dpre_z_partial,
                                                                     dh, R * D);
    DPC_LAUNCH_CHECK();
    return DPC_OK;
}

extern "C" int dpc_gru_bwd_zr(const float* dhr, const float* h, const float* r, const float* z,
                              const float* dz_partial, float* dpre_zr, float* dh, int64_t R, int D, void* stream) {
    DPC_REQUIRE(dhr && h && r && z && dz_partial && dpre_zr && dh && R > 0 && D > 0, "dpc_gru_bwd_zr: bad args");
    gru_bwd_zr_kernel<<<ew_grid(R * D), 256, 0, as_stream(stream)>>>(dhr, h, r, z, dz_partial, dpre_zr, dh, R, D);
    DPC_LAUNCH_CHECK();
    return DPC_OK;
}

extern "C" int dpc_bias_relu(const float* x, const float* b, float* y, int relu, int64_t R, int D, void* stream) {
    DPC_REQUIRE(x && y && R > 0 && D > 0, "dpc_bias_relu: bad args");
    bias_relu_kernel<<<ew_grid(R * D), 256, 0, as_stream(stream)>>>(x, b, y, relu, R, D);
    DPC_LAUNCH_CHECK();
    return DPC_OK;
}

extern "C" int dpc_relu_bwd(const float* y, const float* dy, float* dx, int accumulate, int64_t n, void* stream) {
    DPC_REQUIRE(y && dy && dx && n > 0, "dpc_relu_bwd: bad args");
    relu_bwd_kernel<<<ew_grid(n), 256, 0, as_stream(stream)>>>(y, dy, dx, accumulate, n);
    DPC_LAUNCH_CHECK();
    return DPC_OK;
}

extern "C" int dpc_adam_step(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1,
                             float beta2, float eps, float wd, int step, float gscale, void* stream) {
    DPC_REQUIRE(p && g && m && v && n > 0 && step >= 1, "dpc_adam_step: bad args");
    float bc1 = (float)(1.0 - pow((double)beta1, (double)step));
    float bc2 = (float)(1.0 - pow((double)beta2, (double)step));
    adam_kernel<<<ew_grid(n), 256, 0, as_stream(stream)>>>(p, g, m, v, n, lr, beta1, beta2, eps, wd, bc1,
                                                          sqrtf(bc2), gscale);
    DPC_LAUNCH_CHECK();
    return DPC_OK;
}

// ---- pieces used by the LC classifier (eval/model_3d_lc.py): running BatchNorm statistics, ReLU-then-average
// pooling, stand-alone dropout -----------------------------------------------------------------------------
namespace {

__global__ void bn_running_update_kernel(const float* __restrict__ mean, const float* __restrict__ rstd, double n,
                                         float eps, float momentum, float* __restrict__ rm, float* __restrict__ rv, int C) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    const double r = (double)rstd[c];
    double var_b = 1.0 / (r * r) - (double)eps;             // biased batch variance
    if (var_b < 0.0) var_b = 0.0;
    const double var_u = n > 1.0 ? var_b * n / (n - 1.0) : var_b;
    rm[c] = (1.f - momentum) * rm[c] + momentum * mean[c];
    rv[c] = (1.f - momentum) * rv[c] + momentum * (float)var_u;
}

__global__ void rstd_from_var_kernel(const float* __restrict__ var, float eps, float* __restrict__ rstd, int C) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c < C) rstd[c] = (float)(1.0 / sqrt((double)var[c] + (double)eps));
}

// feat[n, e] = mean_t relu(z[n, t, e])
__global__ void relu_pool_fwd_kernel(const float4* __restrict__ z, float4* __restrict__ feat, long long NB, int T, long long E4) {
    const float inv = 1.f / (float)T;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < NB * E4; i += (long long)gridDim.x * blockDim.x) {
        const long long n = i / E4, e = i % E4;
        float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int t = 0; t < T; ++t) {
            const float4 v = z[(n * T + t) * E4 + e];
            s.x += fmaxf(v.x, 0.f); s.y += fmaxf(v.y, 0.f); s.z += fmaxf(v.z, 0.f); s.w += fmaxf(v.w, 0.f);
        }
        feat[i] = make_float4(s.x * inv, s.y * inv, s.z * inv, s.w * inv);
    }
}
__global__ void relu_pool_bwd_kernel(const float4* __restrict__ z, const float4* __restrict__ dfeat, float4* __restrict__ dz,
                                     long long NB, int T, long long E4) {
    const float inv = 1.f / (float)T;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < NB * E4; i += (long long)gridDim.x * blockDim.x) {
        const long long n = i / E4, e = i % E4;
        const float4 d = dfeat[i];
        for (int t = 0; t < T; ++t) {
            const float4 v = z[(n * T + t) * E4 + e];
            dz[(n * T + t) * E4 + e] = make_float4(v.x > 0.f ? d.x * inv : 0.f, v.y > 0.f ? d.y * inv : 0.f,
                                                   v.z > 0.f ? d.z * inv : 0.f, v.w > 0.f ? d.w * inv : 0.f);
        }
    }
}

__global__ void dropout_kernel(const float* __restrict__ x, float* __restrict__ y, float* __restrict__ keep, float p,
                               uint64_t seed, uint64_t offset, long long n) {
    const float scale = 1.f / (1.f - p);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float k = uniform01(seed, offset + (uint64_t)i) >= p ? scale : 0.f;
        keep[i] = k;
        y[i] = x[i] * k;
    }
}
__global__ void mul_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ out, long long n) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        out[i] = a[i] * b[i];
}

}  // namespace

extern "C" int dpc_bn_running_update(const float* mean, const float* rstd, int64_t rows, float eps, float momentum,
                                     float* running_mean, float* running_var, int C, void* stream) {
    DPC_REQUIRE(mean && rstd && running_mean && running_var && rows > 0 && C > 0, "dpc_bn_running_update: bad args");
    bn_running_update_kernel<<<ceil_div(C, 128), 128, 0, as_stream(stream)>>>(mean, rstd, (double)rows, eps, momentum,
                                                                             running_mean, running_var, C);
    DPC_LAUNCH_CHECK();
    return DPC_OK;
}
extern "C" int dpc_bn_rstd_from_var(const float* var, float eps, float* rstd, int C, void* stream) {
    DPC_REQUIRE(var && rstd && C > 0, "dpc_bn_rstd_from_var: bad args");
    rstd_from_var_kernel<<<ceil_div(C, 128), 128, 0, as_stream(stream)>>>(var, eps, rstd, C);
    DPC_LAUNCH_CHECK();
    return DPC_OK;
}
extern "C" int dpc_relu_pool_fwd(const float* z, float* feat, int NB, int T, int64_t E, void* stream) {
    DPC_REQUIRE(z && feat && NB > 0 && T > 0 && E > 0 && E % 4 == 0, "dpc_relu_pool_fwd: bad args");
    relu_pool_fwd_kernel<<<ew_grid((long long)NB * E / 4), 256, 0, as_stream(stream)>>>((const float4*)z, (float4*)feat, NB, T, E / 4);
    DPC_LAUNCH_CHECK();
    return DPC_OK;
}
extern "C" int dpc_relu_pool_bwd(const float* z, const float* dfeat, float* dz, int NB, int T, int64_t E, void* stream) {
    DPC_REQUIRE(z && dfeat && dz && NB > 0 && T > 0 && E > 0 && E % 4 == 0, "dpc_relu_pool_bwd: bad args");
    relu_pool_bwd_kernel<<<ew_grid((long long)NB * E / 4), 256, 0, as_stream(stream)>>>((const float4*)z, (const float4*)dfeat,
                                                                                        (float4*)dz, NB, T, E / 4);
    DPC_LAUNCH_CHECK();
    return DPC_OK;
}
extern "C" int dpc_dropout_fwd(const float* x, float* y, float* keep, float p, uint64_t seed, uint64_t offset, int64_t n,
                               void* stream) {
    DPC_REQUIRE(x && y && keep && n > 0 && p >= 0.f && p < 1.f, "dpc_dropout_fwd: bad args");
    dropout_kernel<<<ew_grid(n), 256, 0, as_stream(stream)>>>(x, y, keep, p, seed, offset, n);
    DPC_LAUNCH_CHECK();
    return DPC_OK;
}
extern "C" int dpc_mul(const float* a, const float* b, float* out, int64_t n, void* stream) {
    DPC_REQUIRE(a && b && out && n > 0, "dpc_mul: bad args");
    mul_kernel<<<ew_grid(n), 256, 0, as_stream(stream)>>>(a, b, out, n);
    DPC_LAUNCH_CHECK();
    return DPC_OK;
}

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

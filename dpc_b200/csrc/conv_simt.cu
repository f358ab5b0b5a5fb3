// Exact-fp32 CUDA-core implicit-GEMM Conv3d (forward, dgrad, wgrad) over channels-last rows, plus
// weight packing.  This is the reference-precision path of the library: it backs every conv shape
// (any stride / padding / kernel) and is what the tcgen05 kernels are validated against.
// Replaces nn.Conv3d at backbone/resnet_2d3d.py:13-31,241-244.
#include "common.cuh"

namespace {

struct ConvP {
    int NB, Ti, Hi, Wi, Ci, To, Ho, Wo, Co;
    int kT, kH, kW, sT, sH, sW, pT, pH, pW;
    long long M;      // rows of the GEMM (output positions for fwd, input positions for dgrad)
    int Ksrc;         // channels of the gathered tensor
    int N;            // channels of the produced tensor
};

constexpr int BM = 128, BN = 64, BK = 16;

// One kernel for forward (DGRAD=false: gather x at strided taps) and dgrad (DGRAD=true: gather dy
// at the taps that map onto each input position).  w is [tap][Ksrc][N] (N contiguous).
template <bool DGRAD>
__global__ void __launch_bounds__(256) conv_gemm_kernel(ConvP p, const float* __restrict__ src,
                                                         const float* __restrict__ w,
                                                         float* __restrict__ dst, int accumulate) {
    __shared__ __align__(16) float As[BK][BM + 4];
    __shared__ __align__(16) float Bs[BK][BN];
    const int tid = threadIdx.x;
    const long long m0 = (long long)blockIdx.x * BM;
    const int n0 = blockIdx.y * BN;
    const int tm = (tid / 16) * 8, tn = (tid % 16) * 4;

    // the two A rows this thread gathers
    const int arow = tid / 4;          // 0..63 (+64)
    const int akq = (tid % 4) * 4;     // k offset within the chunk
    int rn[2], rt[2], rh[2], rw[2];
    bool rvalid[2];
    const int Tr = DGRAD ? p.Ti : p.To, Hr = DGRAD ? p.Hi : p.Ho, Wr = DGRAD ? p.Wi : p.Wo;
#pragma unroll
    for (int e = 0; e < 2; ++e) {
        long long m = m0 + arow + e * 64;
        rvalid[e] = m < p.M;
        long long mm = rvalid[e] ? m : 0;
        rw[e] = (int)(mm % Wr); mm /= Wr;
        rh[e] = (int)(mm % Hr); mm /= Hr;
        rt[e] = (int)(mm % Tr); mm /= Tr;
        rn[e] = (int)mm;
    }
    const int bk = tid / 16, bn4 = (tid % 16) * 4;

    float acc[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    const int taps = p.kT * p.kH * p.kW;
    for (int tap = 0; tap < taps; ++tap) {
        const int kt = tap / (p.kH * p.kW), kh = (tap / p.kW) % p.kH, kw = tap % p.kW;
        const float* aptr[2];
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            bool ok = rvalid[e];
            long long idx = 0;
            if (!DGRAD) {
                int ti = rt[e] * p.sT - p.pT + kt, hi = rh[e] * p.sH - p.pH + kh, wi = rw[e] * p.sW - p.pW + kw;
                ok = ok && ti >= 0 && ti < p.Ti && hi >= 0 && hi < p.Hi && wi >= 0 && wi < p.Wi;
                idx = (((long long)rn[e] * p.Ti + ti) * p.Hi + hi) * p.Wi + wi;
            } else {
                int tt = rt[e] + p.pT - kt, hh = rh[e] + p.pH - kh, ww = rw[e] + p.pW - kw;
                ok = ok && tt >= 0 && hh >= 0 && ww >= 0 && (tt % p.sT) == 0 && (hh % p.sH) == 0 && (ww % p.sW) == 0;
                int to = tt / p.sT, ho = hh / p.sH, wo = ww / p.sW;
                ok = ok && to < p.To && ho < p.Ho && wo < p.Wo;
                idx = (((long long)rn[e] * p.To + to) * p.Ho + ho) * p.Wo + wo;
            }
            aptr[e] = ok ? src + idx * p.Ksrc + akq : nullptr;
        }
        const float* wtap = w + (size_t)tap * p.Ksrc * p.N + n0 + bn4;
        for (int c0 = 0; c0 < p.Ksrc; c0 += BK) {
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (aptr[e]) v = *reinterpret_cast<const float4*>(aptr[e] + c0);
                int r = arow + e * 64;
                As[akq + 0][r] = v.x; As[akq + 1][r] = v.y; As[akq + 2][r] = v.z; As[akq + 3][r] = v.w;
            }
            *reinterpret_cast<float4*>(&Bs[bk][bn4]) =
                *reinterpret_cast<const float4*>(wtap + (size_t)(c0 + bk) * p.N);
            __syncthreads();
#pragma unroll
            for (int k = 0; k < BK; ++k) {
                float4 a0 = *reinterpret_cast<const float4*>(&As[k][tm]);
                float4 a1 = *reinterpret_cast<const float4*>(&As[k][tm + 4]);
                float4 b = *reinterpret_cast<const float4*>(&Bs[k][tn]);
                float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
                float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
                for (int i = 0; i < 8; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
            }
            __syncthreads();
        }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        long long m = m0 + tm + i;
        if (m >= p.M) continue;
        float4* o = reinterpret_cast<float4*>(dst + m * p.N + n0 + tn);
        float4 v = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
        if (accumulate) {
            float4 c = *o;
            v.x += c.x; v.y += c.y; v.z += c.z; v.w += c.w;
        }
        *o = v;
    }
}

// wgrad: dwp[tap][ci][co] = sum_m x[pos(m,tap)][ci] * dy[m][co].  64x64 tile, split over m, atomics.
constexpr int WM = 64, WN = 64, WK = 16;
__global__ void __launch_bounds__(256) conv_wgrad_kernel(ConvP p, const float* __restrict__ x,
                                                          const float* __restrict__ dy,
                                                          float* __restrict__ dwp, long long rows_per_split) {
    __shared__ __align__(16) float As[WK][WM];
    __shared__ __align__(16) float Bs[WK][WN];
    const int tid = threadIdx.x;
    const int ci_tiles = p.Ci / WM;
    const int ci0 = (blockIdx.x % ci_tiles) * WM, co0 = (blockIdx.x / ci_tiles) * WN;
    const int tap = blockIdx.y;
    const int kt = tap / (p.kH * p.kW), kh = (tap / p.kW) % p.kH, kw = tap % p.kW;
    const long long mbeg = (long long)blockIdx.z * rows_per_split;
    long long mend = mbeg + rows_per_split;
    if (mend > p.M) mend = p.M;
    const int tm = (tid / 16) * 4, tn = (tid % 16) * 4;
    const int lk = tid / 16, lc4 = (tid % 16) * 4;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    for (long long mc = mbeg; mc < mend; mc += WK) {
        long long m = mc + lk;
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
        if (m < mend) {
            long long mm = m;
            int wo = (int)(mm % p.Wo); mm /= p.Wo;
            int ho = (int)(mm % p.Ho); mm /= p.Ho;
            int to = (int)(mm % p.To); mm /= p.To;
            int n = (int)mm;
            int ti = to * p.sT - p.pT + kt, hi = ho * p.sH - p.pH + kh, wi = wo * p.sW - p.pW + kw;
            if (ti >= 0 && ti < p.Ti && hi >= 0 && hi < p.Hi && wi >= 0 && wi < p.Wi) {
                long long idx = (((long long)n * p.Ti + ti) * p.Hi + hi) * p.Wi + wi;
                a = *reinterpret_cast<const float4*>(x + idx * p.Ci + ci0 + lc4);
            }
            b = *reinterpret_cast<const float4*>(dy + m * p.Co + co0 + lc4);
        }
        *reinterpret_cast<float4*>(&As[lk][lc4]) = a;
        *reinterpret_cast<float4*>(&Bs[lk][lc4]) = b;
        __syncthreads();
#pragma unroll
        for (int k = 0; k < WK; ++k) {
            float4 av4 = *reinterpret_cast<const float4*>(&As[k][tm]);
            float4 bv4 = *reinterpret_cast<const float4*>(&Bs[k][tn]);
            float av[4] = {av4.x, av4.y, av4.z, av4.w}, bv[4] = {bv4.x, bv4.y, bv4.z, bv4.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        __syncthreads();
    }
    float* o = dwp + ((size_t)tap * p.Ci + ci0 + tm) * p.Co + co0 + tn;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) atomicAdd(o + (size_t)i * p.Co + j, acc[i][j]);
}

__global__ void pack_weight_kernel(const float* __restrict__ w, float* __restrict__ wf,
                                   float* __restrict__ wd, int Co, int Ci, int taps) {
    long long total = (long long)Co * Ci * taps;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        // i indexes wf: [tap][ci][co]
        int co = (int)(i % Co);
        int ci = (int)((i / Co) % Ci);
        int tap = (int)(i / ((long long)Co * Ci));
        float v = w[((size_t)co * Ci + ci) * taps + tap];
        if (wf) wf[i] = v;
        if (wd) wd[((size_t)tap * Co + co) * Ci + ci] = v;
    }
}

__global__ void unpack_wgrad_kernel(const float* __restrict__ dwp, float* __restrict__ dw, int Co,
                                    int Ci, int taps) {
    long long total = (long long)Co * Ci * taps;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        int co = (int)(i % Co);
        int ci = (int)((i / Co) % Ci);
        int tap = (int)(i / ((long long)Co * Ci));
        dw[((size_t)co * Ci + ci) * taps + tap] = dwp[i];
    }
}

int fill_params(const dpc_conv_geom* g, ConvP& p, const char* who) {
    DPC_REQUIRE(g != nullptr, "%s: null geometry", who);
    p.NB = g->NB; p.Ti = g->Ti; p.Hi = g->Hi; p.Wi = g->Wi; p.Ci = g->Ci;
    p.To = g->To; p.Ho = g->Ho; p.Wo = g->Wo; p.Co = g->Co;
    p.kT = g->kT; p.kH = g->kH; p.kW = g->kW; p.sT = g->sT; p.sH = g->sH; p.sW = g->sW;
    p.pT = g->pT; p.pH = g->pH; p.pW = g->pW;
    DPC_REQUIRE(p.NB > 0 && p.Ti > 0 && p.Hi > 0 && p.Wi > 0 && p.To > 0 && p.Ho > 0 && p.Wo > 0,
                "%s: empty tensor", who);
    DPC_REQUIRE(p.Ci % 16 == 0 && p.Co % 16 == 0, "%s: Ci (%d) and Co (%d) must be multiples of 16", who, p.Ci, p.Co);
    DPC_REQUIRE((p.Ti + 2 * p.pT - p.kT) / p.sT + 1 == p.To && (p.Hi + 2 * p.pH - p.kH) / p.sH + 1 == p.Ho &&
                    (p.Wi + 2 * p.pW - p.kW) / p.sW + 1 == p.Wo,
                "%s: output extent does not match the conv arithmetic", who);
    return DPC_OK;
}

}  // namespace

extern "C" int dpc_pack_conv_weight(const float* w, float* wf, float* wd, int Co, int Ci, int taps, void* stream) {
    DPC_REQUIRE(w && (wf || wd) && Co > 0 && Ci > 0 && taps > 0, "dpc_pack_conv_weight: bad args");
    long long total = (long long)Co * Ci * taps;
    int blocks = (int)((total + 255) / 256);
    if (blocks > 148 * 8) blocks = 148 * 8;
    pack_weight_kernel<<<blocks, 256, 0, as_stream(stream)>>>(w, wf, wd, Co, Ci, taps);
    DPC_LAUNCH_CHECK();
    return DPC_OK;
}

extern "C" int dpc_unpack_conv_wgrad(const float* dwp, float* dw, int Co, int Ci, int taps, void* stream) {
    DPC_REQUIRE(dwp && dw && Co > 0 && Ci > 0 && taps > 0, "dpc_unpack_conv_wgrad: bad args");
    long long total = (long long)Co * Ci * taps;
    int blocks = (int)((total + 255) / 256);
    if (blocks > 148 * 8) blocks = 148 * 8;
    unpack_wgrad_kernel<<<blocks, 256, 0, as_stream(stream)>>>(dwp, dw, Co, Ci, taps);
    DPC_LAUNCH_CHECK();
    return DPC_OK;
}

extern "C" int dpc_conv3d_fwd(const dpc_conv_geom* g, const float* x, const float* wf, float* y, void* stream) {
    ConvP p;
    int rc = fill_params(g, p, "dpc_conv3d_fwd");
    if (rc) return rc;
    DPC_REQUIRE(x && wf && y, "dpc_conv3d_fwd: null pointer");
    DPC_REQUIRE(p.Co % BN == 0, "dpc_conv3d_fwd: Co must be a multiple of %d", BN);
    p.M = (long long)p.NB * p.To * p.Ho * p.Wo;
    p.Ksrc = p.Ci;
    p.N = p.Co;
    dim3 grid(ceil_div(p.M, BM), p.Co / BN);
    conv_gemm_kernel<false><<<grid, 256, 0, as_stream(stream)>>>(p, x, wf, y, 0);
    DPC_LAUNCH_CHECK();
    return DPC_OK;
}

extern "C" int dpc_conv3d_dgrad(const dpc_conv_geom* g, const float* dy, const float* wd, float* dx,
                                int accumulate, void* stream) {
    ConvP p;
    int rc = fill_params(g, p, "dpc_conv3d_dgrad");
    if (rc) return rc;
    DPC_REQUIRE(dy && wd && dx, "dpc_conv3d_dgrad: null pointer");
    DPC_REQUIRE(p.Ci % BN == 0, "dpc_conv3d_dgrad: Ci must be a multiple of %d", BN);
    p.M = (long long)p.NB * p.Ti * p.Hi * p.Wi;
    p.Ksrc = p.Co;
    p.N = p.Ci;
    dim3 grid(ceil_div(p.M, BM), p.Ci / BN);
    conv_gemm_kernel<true><<<grid, 256, 0, as_stream(stream)>>>(p, dy, wd, dx, accumulate);
    DPC_LAUNCH_CHECK();
    return DPC_OK;
}

extern "C" int dpc_conv3d_wgrad(const dpc_conv_geom* g, const float* x, const float* dy, float* dwp, void* stream) {
    ConvP p;
    int rc = fill_params(g, p, "dpc_conv3d_wgrad");
    if (rc) return rc;
    DPC_REQUIRE(x && dy && dwp, "dpc_conv3d_wgrad: null pointer");
    DPC_REQUIRE(p.Ci % WM == 0 && p.Co % WN == 0, "dpc_conv3d_wgrad: Ci, Co must be multiples of 64");
    p.M = (long long)p.NB * p.To * p.Ho * p.Wo;
    p.Ksrc = p.Ci;
    p.N = p.Co;
    const int taps = p.kT * p.kH * p.kW;
    const int tiles = (p.Ci / WM) * (p.Co / WN);
    // enough CTAs for ~4 waves, at least 256 rows per split
    long long want = (long long)dpc_num_sms() * 8;
    long long splits = want / ((long long)tiles * taps);
    if (splits < 1) splits = 1;
    long long max_splits = (p.M + 255) / 256;
    if (splits > max_splits) splits = max_splits;
    if (splits > 65535) splits = 65535;
    long long rps = (p.M + splits - 1) / splits;
    rps = ((rps + WK - 1) / WK) * WK;
    splits = (p.M + rps - 1) / rps;
    cudaStream_t st = as_stream(stream);
    DPC_CUDA(cudaMemsetAsync(dwp, 0, sizeof(float) * (size_t)taps * p.Ci * p.Co, st));
    dim3 grid(tiles, taps, (unsigned)splits);
    conv_wgrad_kernel<<<grid, 256, 0, st>>>(p, x, dy, dwp, rps);
    DPC_LAUNCH_CHECK();
    return DPC_OK;
}

// The recurrent head of DPC_RNN.forward as ONE kernel per direction: ConvGRU aggregation over the first N - P blocks,
// then the prediction loop (network_pred: 1x1 conv -> ReLU -> 1x1 conv; ConvGRU step on ReLU(pred)).
//
// With kernel_size 1 every (clip b, position s) row is an independent sequence, so a CTA owns 16 rows for the whole chain:
// the hidden state never leaves the SM between steps, the gate GEMMs ([16 x 512] x [512 x 256] per gate) run on the CUDA
// cores in exact fp32 with the packed weights streamed from L2 (1.5 MB per step, shared by all CTAs), and the sigmoid /
// tanh / blend / dropout are fused behind them.  Round 1 ran this section as ~100 (forward) + ~200 (backward) launches of
// 2048 x 256 x 256 GEMMs and elementwise kernels (3.4 ms of a 77 ms step, launch-latency bound).  The weight gradients are
// NOT reduced here: the backward kernel stores every step's pre-activation gradients, and three tensor-core wgrad GEMMs
// (reduction over all steps x rows) follow.
//
// Replaces ConvGRUCell.forward / ConvGRU.forward (backbone/convrnn.py:24-34,62-88) and the loop at dpc/model_3d.py:62-72.
#include "common.cuh"

namespace {

constexpr int HC_D = 256;         // feature size of the BasicBlock networks (select_backbone.py:7,10)
constexpr int HC_RB = 16;         // rows per CTA
constexpr int HC_T = 256;         // threads = feature columns

__device__ __forceinline__ float hc_sigmoid(float x) { return 1.f / (1.f + expf(-x)); }

// counter-based uniform in [0,1): splitmix64 of (seed, index) -- the same stream as gru.cu's gru_out_kernel
__device__ __forceinline__ float hc_uniform01(uint64_t seed, uint64_t idx) {
    uint64_t z = seed + 0x9E3779B97F4A7C15ull * (idx + 1);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z = z ^ (z >> 31);
    return (float)(z >> 40) * (1.0f / 16777216.0f);
}

// gate weights [D][2D] (x | h columns) -> transposed, gates side by side:  wt_zr [2D][2D] (z | r),  wt_o [2D][D];
// network_pred weights [D][D] -> transposed w0t, w2t [D][D]
__global__ void hc_pack_kernel(const float* __restrict__ Wz, const float* __restrict__ Wr, const float* __restrict__ Wo,
                               const float* __restrict__ W0, const float* __restrict__ W2, float* __restrict__ wt_zr,
                               float* __restrict__ wt_o, float* __restrict__ w0t, float* __restrict__ w2t) {
    const int D = HC_D;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;            // over 2D x D
    if (i >= 2 * D * D) return;
    const int k = i / D, j = i - k * D;
    wt_zr[(size_t)k * 2 * D + j] = Wz[(size_t)j * 2 * D + k];
    wt_zr[(size_t)k * 2 * D + D + j] = Wr[(size_t)j * 2 * D + k];
    wt_o[(size_t)k * D + j] = Wo[(size_t)j * 2 * D + k];
    if (k < D) {
        w0t[(size_t)k * D + j] = W0[(size_t)j * D + k];
        w2t[(size_t)k * D + j] = W2[(size_t)j * D + k];
    }
}

// acc[r] += sum_k in_s[r][k] * w[k][col]  for k in [0, K), one weight column per thread (stride ldw floats between k's)
template <int NG>
__device__ __forceinline__ void hc_gemv16(float (&acc)[NG][HC_RB], const float (*in_s)[2 * HC_D + 4], int K,
                                          const float* __restrict__ w, int ldw, int gstride) {
    // w points at this thread's column of gate 0; gate g's column is gstride floats further
    float wn[NG][4];
#pragma unroll
    for (int g = 0; g < NG; ++g)
#pragma unroll
        for (int e = 0; e < 4; ++e) wn[g][e] = __ldg(w + (size_t)e * ldw + g * gstride);
    for (int k = 0; k < K; k += 4) {
        float wc[NG][4];
#pragma unroll
        for (int g = 0; g < NG; ++g)
#pragma unroll
            for (int e = 0; e < 4; ++e) wc[g][e] = wn[g][e];
        if (k + 4 < K) {                                              // next k-group in flight while this one is used
#pragma unroll
            for (int g = 0; g < NG; ++g)
#pragma unroll
                for (int e = 0; e < 4; ++e) wn[g][e] = __ldg(w + (size_t)(k + 4 + e) * ldw + g * gstride);
        }
#pragma unroll
        for (int r = 0; r < HC_RB; ++r) {
            const float4 v = *reinterpret_cast<const float4*>(&in_s[r][k]);
#pragma unroll
            for (int g = 0; g < NG; ++g)
                acc[g][r] = fmaf(v.w, wc[g][3], fmaf(v.z, wc[g][2], fmaf(v.y, wc[g][1], fmaf(v.x, wc[g][0], acc[g][r]))));
        }
    }
}

struct HcArgs {
    // geometry: rows r = b * S + s, R = B * S; blocks per clip N; aggregate steps Tagg, prediction steps P
    int B, N, S, Tagg, P, R;
    const float* feat;           // [B*N*S, D] ReLU'd pooled features: x_t of row (b, s) = feat[(b*N + t)*S + s]
    const float *wt_zr, *wt_o, *w0t, *w2t;          // packed (forward) ...
    const float *Wz, *Wr, *Wo, *W0, *W2;            // ... and original layouts (backward)
    const float *bz, *br, *bo, *b0, *b2;
    float p_drop;
    unsigned long long seed;
    // saved by the forward for the backward, step-major [T7 = Tagg + P - 1][R][...]
    float* XH;                   // [T7][R][2D]  gate input  [x | h]
    float* XO;                   // [T7][R][2D]  out-gate input  [x | h * r]
    float *Z, *Rg, *O, *Keep;    // [T7][R][D]   (Keep may be null when p_drop == 0)
    float *U, *Hp;               // [P][R][D]    network_pred hidden (post-ReLU) and the state it was predicted from
    float *Pp;                   // [P][R][D]    predictions (pre-ReLU)
    float* pred_rows;            // [B*P*S, D]   predictions in score-row order (b, i, s)
    // backward
    const float* dpred_rows;     // [B*P*S, D]
    float* dfeat;                // [B*N*S, D]   only the Tagg aggregated blocks are written
    float* DZR;                  // [T7][R][2D]  d pre-activation of (z | r)
    float* DO;                   // [T7][R][D]   d pre-activation of the out gate
    float *DP, *DU;              // [P][R][D]    d prediction, d network_pred hidden (pre-ReLU)
};

__global__ void __launch_bounds__(HC_T, 1) head_chain_fwd_kernel(const HcArgs a) {
    const int D = HC_D;
    __shared__ __align__(16) float in_s[HC_RB][2 * HC_D + 4];
    const int j = threadIdx.x;
    const int row0 = blockIdx.x * HC_RB;
    const int T7 = a.Tagg + a.P - 1;
    int fb[HC_RB];                                 // feat row of (b, s) at block 0, or -1 past the end
#pragma unroll
    for (int r = 0; r < HC_RB; ++r) {
        const int row = row0 + r;
        fb[r] = row < a.R ? ((row / a.S) * a.N * a.S + (row % a.S)) : -1;
    }
    const float bz = a.bz[j], br = a.br[j], bo = a.bo[j], b0 = a.b0[j], b2 = a.b2[j];
    const float scale = a.p_drop > 0.f ? 1.f / (1.f - a.p_drop) : 1.f;
    float h[HC_RB], xin[HC_RB];
#pragma unroll
    for (int r = 0; r < HC_RB; ++r) { h[r] = 0.f; xin[r] = 0.f; }
    for (int t = 0; t < T7; ++t) {
        // ---- gate input [x | h] ----
#pragma unroll
        for (int r = 0; r < HC_RB; ++r) {
            float x = xin[r];                                                    // ReLU(prediction) in the prediction loop
            if (t < a.Tagg) x = fb[r] >= 0 ? a.feat[((size_t)fb[r] + (size_t)t * a.S) * D + j] : 0.f;
            in_s[r][j] = x;
            in_s[r][D + j] = h[r];
            if (fb[r] >= 0) {
                float* xh = a.XH + ((size_t)t * a.R + row0 + r) * 2 * D;
                xh[j] = x;
                xh[D + j] = h[r];
            }
        }
        __syncthreads();
        float zr[2][HC_RB];
#pragma unroll
        for (int r = 0; r < HC_RB; ++r) { zr[0][r] = bz; zr[1][r] = br; }
        hc_gemv16<2>(zr, in_s, 2 * D, a.wt_zr + j, 2 * D, D);
#pragma unroll
        for (int r = 0; r < HC_RB; ++r) { zr[0][r] = hc_sigmoid(zr[0][r]); zr[1][r] = hc_sigmoid(zr[1][r]); }
        __syncthreads();                                                         // all reads of the h half are done
#pragma unroll
        for (int r = 0; r < HC_RB; ++r) in_s[r][D + j] = h[r] * zr[1][r];
        __syncthreads();
        float oo[1][HC_RB];
#pragma unroll
        for (int r = 0; r < HC_RB; ++r) oo[0][r] = bo;
        hc_gemv16<1>(oo, in_s, 2 * D, a.wt_o + j, D, 0);
#pragma unroll
        for (int r = 0; r < HC_RB; ++r) {
            const int row = row0 + r;
            const float o = tanhf(oo[0][r]), z = zr[0][r];
            const float hn = h[r] * (1.f - z) + o * z;                           // convrnn.py:33
            float k = 1.f;
            if (a.p_drop > 0.f)
                k = hc_uniform01(a.seed, ((uint64_t)t * a.R + row) * D + j) >= a.p_drop ? scale : 0.f;   // convrnn.py:78
            if (fb[r] >= 0) {
                const size_t e = ((size_t)t * a.R + row) * D + j;
                float* xo = a.XO + ((size_t)t * a.R + row) * 2 * D;
                xo[j] = in_s[r][j];
                xo[D + j] = h[r] * zr[1][r];
                a.Z[e] = z; a.Rg[e] = zr[1][r]; a.O[e] = o;
                if (a.Keep) a.Keep[e] = k;
            }
            h[r] = hn * k;
        }
        __syncthreads();
        // ---- prediction from the current state: after the last aggregation step and after every prediction-loop step ----
        if (t >= a.Tagg - 1) {
            const int i = t - (a.Tagg - 1);
#pragma unroll
            for (int r = 0; r < HC_RB; ++r) in_s[r][j] = h[r];
            __syncthreads();
            float u[1][HC_RB];
#pragma unroll
            for (int r = 0; r < HC_RB; ++r) u[0][r] = b0;
            hc_gemv16<1>(u, in_s, D, a.w0t + j, D, 0);
            __syncthreads();
#pragma unroll
            for (int r = 0; r < HC_RB; ++r) {
                u[0][r] = fmaxf(u[0][r], 0.f);
                in_s[r][j] = u[0][r];
                if (fb[r] >= 0) {
                    const size_t e = ((size_t)i * a.R + row0 + r) * D + j;
                    a.U[e] = u[0][r];
                    a.Hp[e] = h[r];
                }
            }
            __syncthreads();
            float pr[1][HC_RB];
#pragma unroll
            for (int r = 0; r < HC_RB; ++r) pr[0][r] = b2;
            hc_gemv16<1>(pr, in_s, D, a.w2t + j, D, 0);
#pragma unroll
            for (int r = 0; r < HC_RB; ++r) {
                const int row = row0 + r;
                if (fb[r] >= 0) {
                    a.Pp[((size_t)i * a.R + row) * D + j] = pr[0][r];
                    a.pred_rows[(((size_t)(row / a.S) * a.P + i) * a.S + (row % a.S)) * D + j] = pr[0][r];
                }
                xin[r] = fmaxf(pr[0][r], 0.f);                                    // model_3d.py:70: agg(relu(p_tmp), hidden)
            }
            __syncthreads();
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// backward chain: the same rows, steps in reverse.  Per GRU step (h, z, r, o, keep, hr from the forward):
//   d = dh_out * keep;  d_o = d * z;  d_z = d * (o - h);  dh = d * (1 - z)
//   dpo = d_o * (1 - o^2);  dpz = d_z * z (1 - z)
//   [dx | dhr] = dpo . Wo;   dh += dhr * r;   dpr = dhr * h * r (1 - r)
//   [dx | dh] += [dpz | dpr] . [Wz ; Wr]
// ---------------------------------------------------------------------------------------------------------------------
// out[r] (two columns per thread: col and D + col) += sum_j A_s[r][j] * W[j][col (+D)],  W row-major [J][2D]
__device__ __forceinline__ void hc_gemv16_t(float (&acc)[2][HC_RB], const float (*A_s)[2 * HC_D + 4], int J,
                                            const float* __restrict__ W0, const float* __restrict__ W1, int J0) {
    // rows j < J0 come from W0 (row j), rows j >= J0 from W1 (row j - J0); both [.][2D]
    const int D = HC_D;
    for (int j4 = 0; j4 < J; j4 += 4) {
        float wa[4], wb[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int jj = j4 + e;
            const float* wrow = jj < J0 ? W0 + (size_t)jj * 2 * D : W1 + (size_t)(jj - J0) * 2 * D;
            wa[e] = __ldg(wrow + threadIdx.x);
            wb[e] = __ldg(wrow + D + threadIdx.x);
        }
#pragma unroll
        for (int r = 0; r < HC_RB; ++r) {
            const float4 v = *reinterpret_cast<const float4*>(&A_s[r][j4]);
            acc[0][r] = fmaf(v.w, wa[3], fmaf(v.z, wa[2], fmaf(v.y, wa[1], fmaf(v.x, wa[0], acc[0][r]))));
            acc[1][r] = fmaf(v.w, wb[3], fmaf(v.z, wb[2], fmaf(v.y, wb[1], fmaf(v.x, wb[0], acc[1][r]))));
        }
    }
}
// out[r] (one column per thread) += sum_j A_s[r][j] * W[j][col],  W row-major [J][D]
__device__ __forceinline__ void hc_gemv16_t1(float (&acc)[HC_RB], const float (*A_s)[2 * HC_D + 4], int J,
                                             const float* __restrict__ W) {
    const int D = HC_D;
    for (int j4 = 0; j4 < J; j4 += 4) {
        float wa[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) wa[e] = __ldg(W + (size_t)(j4 + e) * D + threadIdx.x);
#pragma unroll
        for (int r = 0; r < HC_RB; ++r) {
            const float4 v = *reinterpret_cast<const float4*>(&A_s[r][j4]);
            acc[r] = fmaf(v.w, wa[3], fmaf(v.z, wa[2], fmaf(v.y, wa[1], fmaf(v.x, wa[0], acc[r]))));
        }
    }
}

__global__ void __launch_bounds__(HC_T, 1) head_chain_bwd_kernel(const HcArgs a) {
    const int D = HC_D;
    __shared__ __align__(16) float A_s[HC_RB][2 * HC_D + 4];
    const int j = threadIdx.x;
    const int row0 = blockIdx.x * HC_RB;
    const int T7 = a.Tagg + a.P - 1;
    bool ok[HC_RB];
#pragma unroll
    for (int r = 0; r < HC_RB; ++r) ok[r] = row0 + r < a.R;
    float dh[HC_RB], dxp[HC_RB];                    // d hidden state; d ReLU(prediction) from the GRU step that consumed it
#pragma unroll
    for (int r = 0; r < HC_RB; ++r) { dh[r] = 0.f; dxp[r] = 0.f; }
    for (int t = T7 - 1; t >= -1; --t) {
        // ---- prediction made from the state AFTER step t (t >= Tagg - 1; the first prediction follows step Tagg - 1) ----
        if (t >= a.Tagg - 1) {
            const int i = t - (a.Tagg - 1);
            float dp[HC_RB];
#pragma unroll
            for (int r = 0; r < HC_RB; ++r) {
                const int row = row0 + r;
                dp[r] = 0.f;
                if (ok[r]) {
                    dp[r] = a.dpred_rows[(((size_t)(row / a.S) * a.P + i) * a.S + (row % a.S)) * D + j];
                    if (i < a.P - 1) dp[r] += a.Pp[((size_t)i * a.R + row) * D + j] > 0.f ? dxp[r] : 0.f;    // through ReLU(p)
                    a.DP[((size_t)i * a.R + row) * D + j] = dp[r];
                }
                A_s[r][j] = dp[r];
            }
            __syncthreads();
            float du[HC_RB];
#pragma unroll
            for (int r = 0; r < HC_RB; ++r) du[r] = 0.f;
            hc_gemv16_t1(du, A_s, D, a.W2);                                       // du = dp . W2   (W2 [D out][D in])
            __syncthreads();
#pragma unroll
            for (int r = 0; r < HC_RB; ++r) {
                const int row = row0 + r;
                const float uu = ok[r] ? a.U[((size_t)i * a.R + row) * D + j] : 0.f;
                du[r] = uu > 0.f ? du[r] : 0.f;
                if (ok[r]) a.DU[((size_t)i * a.R + row) * D + j] = du[r];
                A_s[r][j] = du[r];
            }
            __syncthreads();
            hc_gemv16_t1(dh, A_s, D, a.W0);                                       // dh += du . W0
            __syncthreads();
        }
        if (t < 0) break;
        // ---- GRU step t ----
        float dpz[HC_RB], hh[HC_RB], rr[HC_RB];
#pragma unroll
        for (int r = 0; r < HC_RB; ++r) {
            const int row = row0 + r;
            const size_t e = ((size_t)t * a.R + row) * D + j;
            float z = 0.f, o = 0.f, k = 1.f;
            hh[r] = 0.f; rr[r] = 0.f;
            if (ok[r]) {
                z = a.Z[e]; o = a.O[e]; rr[r] = a.Rg[e];
                hh[r] = a.XH[((size_t)t * a.R + row) * 2 * D + D + j];
                if (a.Keep) k = a.Keep[e];
            }
            const float d = dh[r] * k;
            const float dpo = d * z * (1.f - o * o);
            dpz[r] = d * (o - hh[r]) * z * (1.f - z);
            dh[r] = d * (1.f - z);
            A_s[r][j] = dpo;
            if (ok[r]) a.DO[e] = dpo;
        }
        __syncthreads();
        float g1[2][HC_RB];                            // [dx | dhr] = dpo . Wo
#pragma unroll
        for (int r = 0; r < HC_RB; ++r) { g1[0][r] = 0.f; g1[1][r] = 0.f; }
        hc_gemv16_t(g1, A_s, D, a.Wo, a.Wo, D);
        __syncthreads();
#pragma unroll
        for (int r = 0; r < HC_RB; ++r) {
            const int row = row0 + r;
            const float dhr = g1[1][r];
            const float dpr = dhr * hh[r] * rr[r] * (1.f - rr[r]);
            dh[r] += dhr * rr[r];
            A_s[r][j] = dpz[r];
            A_s[r][D + j] = dpr;
            if (ok[r]) {
                float* dz = a.DZR + ((size_t)t * a.R + row) * 2 * D;
                dz[j] = dpz[r];
                dz[D + j] = dpr;
            }
        }
        __syncthreads();
        float g2[2][HC_RB];                            // [dx | dh] += [dpz | dpr] . [Wz ; Wr]
#pragma unroll
        for (int r = 0; r < HC_RB; ++r) { g2[0][r] = g1[0][r]; g2[1][r] = dh[r]; }
        hc_gemv16_t(g2, A_s, 2 * D, a.Wz, a.Wr, D);
        __syncthreads();
#pragma unroll
        for (int r = 0; r < HC_RB; ++r) {
            const int row = row0 + r;
            dh[r] = g2[1][r];
            if (t < a.Tagg) {
                if (ok[r]) a.dfeat[(((size_t)(row / a.S) * a.N + t) * a.S + (row % a.S)) * D + j] = g2[0][r];
            } else {
                dxp[r] = g2[0][r];                     // gradient of ReLU(prediction t - Tagg), consumed by the next iteration
            }
        }
    }
}

}  // namespace

// pack the ConvGRU / network_pred weights for dpc_head_chain_fwd (wt_zr [2D][2D], wt_o [2D][D], w0t / w2t [D][D]; D = 256)
extern "C" int dpc_head_chain_pack(const float* Wz, const float* Wr, const float* Wo, const float* W0, const float* W2,
                                   float* wt_zr, float* wt_o, float* w0t, float* w2t, void* stream) {
    DPC_REQUIRE(Wz && Wr && Wo && W0 && W2 && wt_zr && wt_o && w0t && w2t, "dpc_head_chain_pack: null pointer");
    hc_pack_kernel<<<(2 * HC_D * HC_D + 255) / 256, 256, 0, as_stream(stream)>>>(Wz, Wr, Wo, W0, W2, wt_zr, wt_o, w0t, w2t);
    DPC_LAUNCH_CHECK();
    return DPC_OK;
}

// ConvGRU aggregation + prediction loop, forward (D = 256, kernel_size 1).  Saved tensors are step-major, see HcArgs.
extern "C" int dpc_head_chain_fwd(const float* feat, const float* wt_zr, const float* wt_o, const float* w0t, const float* w2t,
                                  const float* bz, const float* br, const float* bo, const float* b0, const float* b2,
                                  int B, int N, int S, int P, float p_drop, uint64_t seed, float* XH, float* XO, float* Z,
                                  float* Rg, float* O, float* Keep, float* U, float* Hp, float* Pp, float* pred_rows,
                                  void* stream) {
    DPC_REQUIRE(feat && wt_zr && wt_o && w0t && w2t && bz && br && bo && b0 && b2 && XH && XO && Z && Rg && O && U && Hp && Pp &&
                    pred_rows, "dpc_head_chain_fwd: null pointer");
    DPC_REQUIRE(B > 0 && S > 0 && P > 0 && N > P, "dpc_head_chain_fwd: bad geometry B %d N %d S %d P %d", B, N, S, P);
    DPC_REQUIRE(p_drop >= 0.f && p_drop < 1.f && (p_drop == 0.f || Keep), "dpc_head_chain_fwd: dropout needs the keep buffer");
    HcArgs a;
    memset(&a, 0, sizeof(a));
    a.B = B; a.N = N; a.S = S; a.P = P; a.Tagg = N - P; a.R = B * S;
    a.feat = feat; a.wt_zr = wt_zr; a.wt_o = wt_o; a.w0t = w0t; a.w2t = w2t;
    a.bz = bz; a.br = br; a.bo = bo; a.b0 = b0; a.b2 = b2;
    a.p_drop = p_drop; a.seed = seed;
    a.XH = XH; a.XO = XO; a.Z = Z; a.Rg = Rg; a.O = O; a.Keep = p_drop > 0.f ? Keep : nullptr; a.U = U; a.Hp = Hp; a.Pp = Pp;
    a.pred_rows = pred_rows;
    head_chain_fwd_kernel<<<(a.R + HC_RB - 1) / HC_RB, HC_T, 0, as_stream(stream)>>>(a);
    DPC_LAUNCH_CHECK();
    return DPC_OK;
}

// backward of the chain: dfeat (the Tagg aggregated blocks of [B*N*S, D]; the caller zeroes the rest) and every step's
// pre-activation gradients DZR [T7][R][2D], DO [T7][R][D], DP / DU [P][R][D] for the weight-gradient GEMMs
extern "C" int dpc_head_chain_bwd(const float* dpred_rows, const float* Wz, const float* Wr, const float* Wo, const float* W0,
                                  const float* W2, int B, int N, int S, int P, const float* XH, const float* Z, const float* Rg,
                                  const float* O, const float* Keep, const float* U, const float* Pp, float* dfeat, float* DZR,
                                  float* DO, float* DP, float* DU, void* stream) {
    DPC_REQUIRE(dpred_rows && Wz && Wr && Wo && W0 && W2 && XH && Z && Rg && O && U && Pp && dfeat && DZR && DO && DP && DU,
                "dpc_head_chain_bwd: null pointer");
    DPC_REQUIRE(B > 0 && S > 0 && P > 0 && N > P, "dpc_head_chain_bwd: bad geometry B %d N %d S %d P %d", B, N, S, P);
    HcArgs a;
    memset(&a, 0, sizeof(a));
    a.B = B; a.N = N; a.S = S; a.P = P; a.Tagg = N - P; a.R = B * S;
    a.Wz = Wz; a.Wr = Wr; a.Wo = Wo; a.W0 = W0; a.W2 = W2;
    a.XH = const_cast<float*>(XH); a.Z = const_cast<float*>(Z); a.Rg = const_cast<float*>(Rg); a.O = const_cast<float*>(O);
    a.Keep = const_cast<float*>(Keep); a.U = const_cast<float*>(U); a.Pp = const_cast<float*>(Pp);
    a.dpred_rows = dpred_rows; a.dfeat = dfeat; a.DZR = DZR; a.DO = DO; a.DP = DP; a.DU = DU;
    head_chain_bwd_kernel<<<(a.R + HC_RB - 1) / HC_RB, HC_T, 0, as_stream(stream)>>>(a);
    DPC_LAUNCH_CHECK();
    return DPC_OK;
}

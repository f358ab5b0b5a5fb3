// NCE head: closed-form int8 mask, fused cross-entropy (+ top-1/3/5) over the dense score matrix,
// and its gradient.  Warp-shuffle reductions, one CTA per score row, single pass over HBM.
// Replaces the Python mask loops (dpc/model_3d.py:86-96), process_output / argmax /
// nn.CrossEntropyLoss (dpc/main.py:178-185,213-217) and calc_topk_accuracy (utils/utils.py:38-55).
#include "common.cuh"

namespace {

__global__ void mask_fill_kernel(int8_t* __restrict__ mask, int B, int P, int SQ) {
    // mask[b,p,s,b2,p2,s2]: b==b2 ? (s==s2 ? (p==p2 ? 1 : -1) : -3) : 0
    const long long M = (long long)B * P * SQ;
    const long long total = M * M;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        long long row = i / M, col = i % M;
        int s = (int)(row % SQ), p = (int)((row / SQ) % P), b = (int)(row / ((long long)SQ * P));
        int s2 = (int)(col % SQ), p2 = (int)((col / SQ) % P), b2 = (int)(col / ((long long)SQ * P));
        int8_t v = 0;
        if (b == b2) v = (s == s2) ? ((p == p2) ? 1 : -1) : -3;
        mask[i] = v;
    }
}

// one CTA per row: online softmax (max, sum) + rank of the diagonal entry
// score is [rows, M]; the positive of row i is column i % M (M == rows on one device; under the
// reference's DataParallel gather rows = n_gpu * M, dpc/main.py:212-215)
__global__ void __launch_bounds__(256) ce_fwd_kernel(const float* __restrict__ score, int rows, int M,
                                                      float* __restrict__ lse, float* __restrict__ out) {
    const int row = blockIdx.x;
    const float* s = score + (size_t)row * M;
    const float diag = s[row % M];
    float mx = -INFINITY, sum = 0.f;
    int greater = 0;
    for (int j = threadIdx.x; j < M; j += 256) {
        float v = s[j];
        greater += (v > diag) ? 1 : 0;
        if (v > mx) { sum = sum * expf(mx - v) + 1.f; mx = v; }
        else sum += expf(v - mx);
    }
    // warp reduce (max, sum) pairs
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        float omx = __shfl_xor_sync(0xffffffffu, mx, o);
        float osum = __shfl_xor_sync(0xffffffffu, sum, o);
        int og = __shfl_xor_sync(0xffffffffu, greater, o);
        float nm = fmaxf(mx, omx);
        float a = (mx == -INFINITY) ? 0.f : sum * expf(mx - nm);
        float b = (omx == -INFINITY) ? 0.f : osum * expf(omx - nm);
        sum = a + b; mx = nm; greater += og;
    }
    __shared__ float smx[8], ssum[8];
    __shared__ int sg[8];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) { smx[warp] = mx; ssum[warp] = sum; sg[warp] = greater; }
    __syncthreads();
    if (threadIdx.x == 0) {
        float m = -INFINITY;
        for (int w = 0; w < 8; ++w) m = fmaxf(m, smx[w]);
        float t = 0.f;
        int g = 0;
        for (int w = 0; w < 8; ++w) {
            if (smx[w] != -INFINITY) t += ssum[w] * expf(smx[w] - m);
            g += sg[w];
        }
        float l = m + logf(t);
        lse[row] = l;
        const float inv = 1.f / (float)rows;
        atomicAdd(out + 0, (l - diag) * inv);
        if (g < 1) atomicAdd(out + 1, inv);
        if (g < 3) atomicAdd(out + 2, inv);
        if (g < 5) atomicAdd(out + 3, inv);
    }
}

__global__ void __launch_bounds__(256) ce_bwd_kernel(const float* __restrict__ score, const float* __restrict__ lse,
                                                      const float* __restrict__ gscale, float* __restrict__ dscore,
                                                      int rows, int M) {
    const int row = blockIdx.x;
    const float l = lse[row];
    const float k = (gscale ? gscale[0] : 1.f) / (float)rows;
    const int tgt = row % M;
    const float* s = score + (size_t)row * M;
    float* d = dscore + (size_t)row * M;
    for (int j = threadIdx.x; j < M; j += 256) {
        float p = expf(s[j] - l);
        d[j] = k * (p - (j == tgt ? 1.f : 0.f));
    }
}

}  // namespace

extern "C" int dpc_nce_mask_fill(int8_t* mask, int B, int P, int SQ, void* stream) {
    DPC_REQUIRE(mask && B > 0 && P > 0 && SQ > 0, "dpc_nce_mask_fill: bad args");
    long long M = (long long)B * P * SQ;
    long long blocks = (M * M + 255) / 256;
    long long cap = (long long)dpc_num_sms() * 32;
    mask_fill_kernel<<<(int)(blocks < cap ? blocks : cap), 256, 0, as_stream(stream)>>>(mask, B, P, SQ);
    DPC_LAUNCH_CHECK();
    return DPC_OK;
}

extern "C" int dpc_nce_ce_fwd(const float* score, int rows, int M, float* lse, float* out, void* stream) {
    DPC_REQUIRE(score && lse && out && M > 0 && rows > 0, "dpc_nce_ce_fwd: bad args");
    cudaStream_t st = as_stream(stream);
    DPC_CUDA(cudaMemsetAsync(out, 0, 4 * sizeof(float), st));
    ce_fwd_kernel<<<rows, 256, 0, st>>>(score, rows, M, lse, out);
    DPC_LAUNCH_CHECK();
    return DPC_OK;
}

extern "C" int dpc_nce_ce_bwd(const float* score, const float* lse, const float* gscale, float* dscore,
                              int rows, int M, void* stream) {
    DPC_REQUIRE(score && lse && dscore && M > 0 && rows > 0, "dpc_nce_ce_bwd: bad args");
    ce_bwd_kernel<<<rows, 256, 0, as_stream(stream)>>>(score, lse, gscale, dscore, rows, M);
    DPC_LAUNCH_CHECK();
    return DPC_OK;
}

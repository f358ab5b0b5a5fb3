// Generic fp32 CUDA-core GEMM with arbitrary transposes / leading dimensions.
// Used for the small, latency-bound GEMMs of the ConvGRU / predictor (convrnn.py:29-33,
// model_3d.py:36-40,68) and as the exact-fp32 fallback shape for the score matmul.
#include "common.cuh"

namespace {

constexpr int TM = 64, TN = 64, TK = 16;

// SPLITK: grid.z slices of K; partial products are added to C with fp32 atomics (C pre-scaled by beta)
template <bool TA, bool TB, bool SPLITK>
__global__ void __launch_bounds__(256) gemm_f32_kernel(int M, int N, int K, float alpha,
                                                        const float* __restrict__ A, int lda,
                                                        const float* __restrict__ B, int ldb,
                                                        float beta, float* __restrict__ C, int ldc, int kslice) {
    __shared__ float As[TK][TM + 4];
    __shared__ float Bs[TK][TN + 4];
    const int tid = threadIdx.x;
    const int m0 = blockIdx.y * TM, n0 = blockIdx.x * TN;
    const int tm = (tid / 16) * 4, tn = (tid % 16) * 4;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    const int kbeg = SPLITK ? blockIdx.z * kslice : 0;
    const int kend = SPLITK ? (kbeg + kslice < K ? kbeg + kslice : K) : K;
    for (int k0 = kbeg; k0 < kend; k0 += TK) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            int idx = tid + e * 256;
            int m, k;
            if (TA) { k = idx / TM; m = idx % TM; } else { m = idx / TK; k = idx % TK; }
            int gm = m0 + m, gk = k0 + k;
            float v = 0.f;
            if (gm < M && gk < kend) v = TA ? A[(size_t)gk * lda + gm] : A[(size_t)gm * lda + gk];
            As[k][m] = v;
        }
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            int idx = tid + e * 256;
            int n, k;
            if (TB) { n = idx / TK; k = idx % TK; } else { k = idx / TN; n = idx % TN; }
            int gn = n0 + n, gk = k0 + k;
            float v = 0.f;
            if (gn < N && gk < kend) v = TB ? B[(size_t)gn * ldb + gk] : B[(size_t)gk * ldb + gn];
            Bs[k][n] = v;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < TK; ++k) {
            float4 a = *reinterpret_cast<const float4*>(&As[k][tm]);
            float4 b = *reinterpret_cast<const float4*>(&Bs[k][tn]);
            float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        int gm = m0 + tm + i;
        if (gm >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int gn = n0 + tn + j;
            if (gn >= N) continue;
            float* c = &C[(size_t)gm * ldc + gn];
            float v = alpha * acc[i][j];
            if (SPLITK) { atomicAdd(c, v); continue; }
            if (beta != 0.f) v += beta * (*c);
            *c = v;
        }
    }
}

// Full-tile fast path (M % 64 == N % 64 == 0, K slice % 16 == 0, 16-byte aligned operands): one float4 global load
// per thread and operand per k-tile, issued one k-tile ahead of the FMAs (register prefetch), so the global latency
// of the small latency-bound head GEMMs (2048 x 256 x 256) overlaps the math.
template <bool TA, bool TB, bool SPLITK>
__global__ void __launch_bounds__(256) gemm_f32_vec_kernel(int M, int N, int K, float alpha,
                                                            const float* __restrict__ A, int lda,
                                                            const float* __restrict__ B, int ldb,
                                                            float beta, float* __restrict__ C, int ldc, int kslice) {
    __shared__ __align__(16) float As[TK][TM + 4];
    __shared__ __align__(16) float Bs[TK][TN + 4];
    const int tid = threadIdx.x;
    const int m0 = blockIdx.y * TM, n0 = blockIdx.x * TN;
    const int tm = (tid / 16) * 4, tn = (tid % 16) * 4;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    const int kbeg = SPLITK ? blockIdx.z * kslice : 0;
    const int kend = SPLITK ? (kbeg + kslice < K ? kbeg + kslice : K) : K;
    // this thread's float4 of the A / B k-tile
    const int am = TA ? (tid % 16) * 4 : tid / 4, ak = TA ? tid / 16 : (tid % 4) * 4;
    const int bn = TB ? tid / 4 : (tid % 16) * 4, bk = TB ? (tid % 4) * 4 : tid / 16;
    auto lda4 = [&](int k0) {
        return TA ? *reinterpret_cast<const float4*>(A + (size_t)(k0 + ak) * lda + m0 + am)
                  : *reinterpret_cast<const float4*>(A + (size_t)(m0 + am) * lda + k0 + ak);
    };
    auto ldb4 = [&](int k0) {
        return TB ? *reinterpret_cast<const float4*>(B + (size_t)(n0 + bn) * ldb + k0 + bk)
                  : *reinterpret_cast<const float4*>(B + (size_t)(k0 + bk) * ldb + n0 + bn);
    };
    float4 pa = lda4(kbeg), pb = ldb4(kbeg);
    for (int k0 = kbeg; k0 < kend; k0 += TK) {
        if (TA) *reinterpret_cast<float4*>(&As[ak][am]) = pa;
        else { As[ak][am] = pa.x; As[ak + 1][am] = pa.y; As[ak + 2][am] = pa.z; As[ak + 3][am] = pa.w; }
        if (TB) { Bs[bk][bn] = pb.x; Bs[bk + 1][bn] = pb.y; Bs[bk + 2][bn] = pb.z; Bs[bk + 3][bn] = pb.w; }
        else *reinterpret_cast<float4*>(&Bs[bk][bn]) = pb;
        __syncthreads();
        if (k0 + TK < kend) { pa = lda4(k0 + TK); pb = ldb4(k0 + TK); }
#pragma unroll
        for (int k = 0; k < TK; ++k) {
            const float4 a = *reinterpret_cast<const float4*>(&As[k][tm]);
            const float4 b = *reinterpret_cast<const float4*>(&Bs[k][tn]);
            const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        float* c = &C[(size_t)(m0 + tm + i) * ldc + n0 + tn];
        if (SPLITK) {
#pragma unroll
            for (int j = 0; j < 4; ++j) atomicAdd(c + j, alpha * acc[i][j]);
        } else {
            float4 v = make_float4(alpha * acc[i][0], alpha * acc[i][1], alpha * acc[i][2], alpha * acc[i][3]);
            if (beta != 0.f) {
                const float4 o = *reinterpret_cast<const float4*>(c);
                v.x += beta * o.x; v.y += beta * o.y; v.z += beta * o.z; v.w += beta * o.w;
            }
            *reinterpret_cast<float4*>(c) = v;
        }
    }
}

__global__ void scale_matrix_kernel(float* __restrict__ C, int M, int N, int ldc, float beta) {
    long long total = (long long)M * N;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        float* c = &C[(size_t)(i / N) * ldc + (i % N)];
        *c = beta == 0.f ? 0.f : beta * (*c);
    }
}

__global__ void gather_rows_kernel(const float4* __restrict__ src, float4* __restrict__ dst,
                                   long long rows, int D4, long long inner, long long outer,
                                   long long offset) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long total = rows * D4;
    for (; i < total; i += (long long)gridDim.x * blockDim.x) {
        long long r = i / D4;
        int c = (int)(i % D4);
        long long sr = (r / inner) * outer + offset + r % inner;
        dst[i] = src[sr * D4 + c];
    }
}

__global__ void scatter_rows_kernel(const float4* __restrict__ src, float4* __restrict__ dst,
                                    long long rows, int D4, long long inner, long long outer,
                                    long long offset, int accumulate) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long total = rows * D4;
    for (; i < total; i += (long long)gridDim.x * blockDim.x) {
        long long r = i / D4;
        int c = (int)(i % D4);
        long long dr = (r / inner) * outer + offset + r % inner;
        float4 v = src[i];
        if (accumulate) {
            float4 o = dst[dr * D4 + c];
            v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w;
        }
        dst[dr * D4 + c] = v;
    }
}

// colsum: grid.x over column groups of 32, grid.y over row chunks; partial sums are added with atomics
// (`out` is zeroed by the launcher unless it accumulates)
__global__ void __launch_bounds__(256) colsum_kernel(const float* __restrict__ A, long long rows, int N,
                                                      float* __restrict__ out, long long chunk) {
    __shared__ float red[8][33];
    int col = blockIdx.x * 32 + (threadIdx.x & 31);
    int rl = threadIdx.x >> 5;
    const long long r0 = (long long)blockIdx.y * chunk;
    const long long r1 = r0 + chunk < rows ? r0 + chunk : rows;
    float s = 0.f;
    if (col < N)
        for (long long r = r0 + rl; r < r1; r += 8) s += A[r * N + col];
    red[rl][threadIdx.x & 31] = s;
    __syncthreads();
    if (rl == 0 && col < N) {
        float t = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) t += red[i][threadIdx.x & 31];
        atomicAdd(out + col, t);
    }
}

}  // namespace

extern "C" int dpc_gemm_f32(int transA, int transB, int M, int N, int K, float alpha, const float* A,
                            int lda, const float* B, int ldb, float beta, float* C, int ldc,
                            void* stream) {
    DPC_REQUIRE(M > 0 && N > 0 && K > 0, "dpc_gemm_f32: bad dims %d %d %d", M, N, K);
    DPC_REQUIRE(A && B && C, "dpc_gemm_f32: null pointer");
    dim3 grid(ceil_div(N, TN), ceil_div(M, TM));
    cudaStream_t st = as_stream(stream);
    // few output tiles but a long reduction (the GRU / predictor weight gradients: 256x256 outputs over
    // thousands of rows): slice K across CTAs so the whole chip works on it
    const int tiles = (int)(grid.x * grid.y);
    int splits = 1;
    if (tiles < dpc_num_sms() && K >= 512) {
        splits = (2 * dpc_num_sms() + tiles - 1) / tiles;
        if (splits > K / 128) splits = K / 128;
        if (splits < 1) splits = 1;
    }
    const bool vec = M % TM == 0 && N % TN == 0 && K % TK == 0 && lda % 4 == 0 && ldb % 4 == 0 && ldc % 4 == 0 &&
                     ((uintptr_t)A & 15) == 0 && ((uintptr_t)B & 15) == 0 && ((uintptr_t)C & 15) == 0;
    if (splits > 1) {
        int kslice = ((K + splits - 1) / splits + TK - 1) / TK * TK;
        splits = (K + kslice - 1) / kslice;
        if (beta != 1.f) {
            scale_matrix_kernel<<<ceil_div((long long)M * N, 256), 256, 0, st>>>(C, M, N, ldc, beta);
            DPC_LAUNCH_CHECK();
        }
        grid.z = splits;
#define GEMM_SK(TA_, TB_)                                                                                              \
    do {                                                                                                               \
        if (vec) gemm_f32_vec_kernel<TA_, TB_, true><<<grid, 256, 0, st>>>(M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, kslice); \
        else gemm_f32_kernel<TA_, TB_, true><<<grid, 256, 0, st>>>(M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, kslice);         \
    } while (0)
        if (!transA && !transB) GEMM_SK(false, false);
        else if (!transA && transB) GEMM_SK(false, true);
        else if (transA && !transB) GEMM_SK(true, false);
        else GEMM_SK(true, true);
#undef GEMM_SK
        DPC_LAUNCH_CHECK();
        return DPC_OK;
    }
#define GEMM_1(TA_, TB_)                                                                                               \
    do {                                                                                                               \
        if (vec) gemm_f32_vec_kernel<TA_, TB_, false><<<grid, 256, 0, st>>>(M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, K); \
        else gemm_f32_kernel<TA_, TB_, false><<<grid, 256, 0, st>>>(M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, K);         \
    } while (0)
    if (!transA && !transB) GEMM_1(false, false);
    else if (!transA && transB) GEMM_1(false, true);
    else if (transA && !transB) GEMM_1(true, false);
    else GEMM_1(true, true);
#undef GEMM_1
    DPC_LAUNCH_CHECK();
    return DPC_OK;
}

extern "C" int dpc_gather_rows(const float* src, float* dst, int64_t rows, int D, int64_t inner,
                               int64_t outer, int64_t offset, void* stream) {
    DPC_REQUIRE(D % 4 == 0 && rows > 0 && inner > 0, "dpc_gather_rows: bad args");
    long long total = rows * (D / 4);
    int blocks = (int)((total + 255) / 256);
    if (blocks > 148 * 16) blocks = 148 * 16;
    gather_rows_kernel<<<blocks, 256, 0, as_stream(stream)>>>((const float4*)src, (float4*)dst, rows, D / 4,
                                                             inner, outer, offset);
    DPC_LAUNCH_CHECK();
    return DPC_OK;
}

extern "C" int dpc_scatter_rows(const float* src, float* dst, int64_t rows, int D, int64_t inner,
                                int64_t outer, int64_t offset, int accumulate, void* stream) {
    DPC_REQUIRE(D % 4 == 0 && rows > 0 && inner > 0, "dpc_scatter_rows: bad args");
    long long total = rows * (D / 4);
    int blocks = (int)((total + 255) / 256);
    if (blocks > 148 * 16) blocks = 148 * 16;
    scatter_rows_kernel<<<blocks, 256, 0, as_stream(stream)>>>((const float4*)src, (float4*)dst, rows, D / 4,
                                                              inner, outer, offset, accumulate);
    DPC_LAUNCH_CHECK();
    return DPC_OK;
}

extern "C" int dpc_colsum(const float* A, int64_t rows, int N, float* out, int accumulate, void* stream) {
    DPC_REQUIRE(rows > 0 && N > 0, "dpc_colsum: bad args");
    if (!accumulate) DPC_CUDA(cudaMemsetAsync(out, 0, sizeof(float) * (size_t)N, as_stream(stream)));
    long long chunks = (rows + 127) / 128;                       // >= 128 rows per block
    const long long cap = (2ll * dpc_num_sms() + ceil_div(N, 32) - 1) / ceil_div(N, 32);
    if (chunks > cap) chunks = cap;
    if (chunks < 1) chunks = 1;
    const long long chunk = (rows + chunks - 1) / chunks;
    colsum_kernel<<<dim3(ceil_div(N, 32), (unsigned)((rows + chunk - 1) / chunk)), 256, 0, as_stream(stream)>>>(A, rows, N, out, chunk);
    DPC_LAUNCH_CHECK();
    return DPC_OK;
}

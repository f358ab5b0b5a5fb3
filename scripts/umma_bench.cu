// Microbenchmark: issue rate of tcgen05.mma (kind::f16, bf16 operands from shared memory, cta_group::1) for the
// instruction shapes the conv kernels use.  One CTA per SM, one issuing thread, operands are zeros.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_bench scripts/umma_bench.cu && ./umma_bench
#include "../dpc_b200/csrc/tc_common.cuh"
#include <vector>

void dpc_set_error(const char*, ...) {}
void dpc_count_launch(int) {}

struct Cfg {
    int n[4];        // N of the up-to-4 instructions of one "step" (0 = unused)
    int dcol[4];     // accumulator column of each
    int aoff[4];     // A operand byte offset of each
    int boff[4];     // B operand byte offset of each
    int mn_major;    // 1: both operands MN-major
    int m;           // 128 or 64
    int commit_every; // extra tcgen05.commit (to a barrier nobody waits on) every this many steps (0 = never)
    const char* name;
};

template <int NM>
__global__ void __launch_bounds__(128, 1) umma_bench_kernel(Cfg c, int steps, long long* out) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const uint32_t sbase = smem_u32(smem);
    const uint32_t bar = sbase + 160 * 1024, tptr = bar + 8;
    for (uint32_t i = threadIdx.x; i < 160 * 1024 / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        mbar_init(bar + 16, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (threadIdx.x < 32) tmem_alloc(tptr, 512);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *reinterpret_cast<const uint32_t*>(smem + 160 * 1024 + 8);
    if (threadIdx.x == 0) {
        uint32_t idesc[NM], dc[NM];
        uint64_t ad[NM], bd[NM];
#pragma unroll
        for (int j = 0; j < NM; ++j) {
            idesc[j] = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(c.n[j] >> 3) << 17) | ((uint32_t)(c.m >> 4) << 24);
            dc[j] = tmem + c.dcol[j];
            if (c.mn_major) {
                idesc[j] |= (1u << 15) | (1u << 16);
                ad[j] = make_mnmajor_sw128_desc(sbase + c.aoff[j], 8192);
                bd[j] = make_mnmajor_sw128_desc(sbase + c.boff[j], 8192);
            } else {
                ad[j] = make_kmajor_sw128_desc(sbase + c.aoff[j]);
                bd[j] = make_kmajor_sw128_desc(sbase + c.boff[j]);
            }
        }
        const uint64_t kstep = c.mn_major ? 128 : 2;
        uint32_t phase = 0;
        for (int rep = 0; rep < 2; ++rep) {              // rep 0 = warm-up
            const long long t0 = clock64();
            for (int s = 0; s < steps; s += 4) {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
#pragma unroll
                    for (int j = 0; j < NM; ++j) umma_bf16(dc[j], ad[j] + k * kstep, bd[j] + k * kstep, idesc[j], 1u);
                }
                if (c.commit_every && ((s >> 2) % c.commit_every) == 0) umma_commit(bar + 16);
            }
            umma_commit(bar);
            mbar_wait(bar, phase);
            phase ^= 1u;
            const long long t1 = clock64();
            if (rep == 1 && blockIdx.x == 0) out[0] = t1 - t0;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc(tmem, 512);
}

int main() {
    const int A0 = 0, A1 = 32768, B0 = 65536, B1 = 65536 + 32768;     // 32 KB regions
    std::vector<Cfg> cfgs = {
        {{256, 0, 0, 0}, {0, 0, 0, 0}, {A0, 0, 0, 0}, {B0, 0, 0, 0}, 0, 128, 0, "N=256, one accumulator"},
        {{256, 256, 0, 0}, {0, 256, 0, 0}, {A0, A0, 0, 0}, {B0, B0, 0, 0}, 0, 128, 0, "N=256 x2, two accumulators"},
        {{128, 0, 0, 0}, {0, 0, 0, 0}, {A0, 0, 0, 0}, {B0, 0, 0, 0}, 0, 128, 0, "N=128, one accumulator"},
        {{128, 128, 0, 0}, {0, 128, 0, 0}, {A0, A0, 0, 0}, {B0, B0, 0, 0}, 0, 128, 0, "N=128 x2, two accumulators"},
        {{64, 0, 0, 0}, {0, 0, 0, 0}, {A0, 0, 0, 0}, {B0, 0, 0, 0}, 0, 128, 0, "N=64, one accumulator"},
        {{64, 64, 0, 0}, {0, 64, 0, 0}, {A0, A0, 0, 0}, {B0, B0, 0, 0}, 0, 128, 0, "N=64 x2, two accumulators, same A"},
        {{64, 64, 64, 64}, {0, 64, 128, 192}, {A0, A0, A0, A0}, {B0, B0, B0, B0}, 0, 128, 0, "N=64 x4, four accumulators"},
        {{64, 64, 64, 0}, {0, 64, 64, 0}, {A0, A0, A1, 0}, {B0, B1, B0, 0}, 0, 128, 0, "3xBF16 step, N=64 (old layer1 pattern)"},
        {{128, 64, 0, 0}, {0, 64, 0, 0}, {A0, A1, 0, 0}, {B0, B0, 0, 0}, 0, 128, 0, "wide step N=128 + N=64 (halo pattern)"},
        {{256, 128, 0, 0}, {0, 128, 0, 0}, {A0, A1, 0, 0}, {B0, B0, 0, 0}, 0, 128, 0, "wide step N=256 + N=128 (layer2 pattern)"},
        {{256, 256, 256, 0}, {0, 256, 256, 0}, {A0, A0, A1, 0}, {B0, B1, B0, 0}, 0, 128, 0, "3xBF16 step, N=256 (layer3 pattern)"},
        {{64, 64, 64, 0}, {0, 0, 0, 0}, {A0, A0, A1, 0}, {B0, B1, B0, 0}, 1, 128, 0, "MN-major 3 x N=64 into one accumulator (wgrad-halo pattern)"},
        {{256, 256, 256, 0}, {0, 256, 256, 0}, {A0, A0, A1, 0}, {B0, B1, B0, 0}, 1, 128, 0, "MN-major 3xBF16 step N=256 (wgrad pattern)"},
        {{64, 0, 0, 0}, {0, 0, 0, 0}, {A0, 0, 0, 0}, {B0, 0, 0, 0}, 0, 64, 0, "M=64 N=64"},
        {{256, 0, 0, 0}, {0, 0, 0, 0}, {A0, 0, 0, 0}, {B0, 0, 0, 0}, 0, 64, 0, "M=64 N=256"},
        {{128, 64, 0, 0}, {0, 64, 0, 0}, {A0 + 384, A1 + 384, 0, 0}, {B0, B0, 0, 0}, 0, 128, 0, "halo pattern, A start shifted by 3 rows"},
        {{128, 64, 0, 0}, {0, 64, 0, 0}, {A0 + 640, A1 + 640, 0, 0}, {B0, B0, 0, 0}, 0, 128, 0, "halo pattern, A start shifted by 5 rows"},
        {{128, 64, 0, 0}, {0, 64, 0, 0}, {A0, A1, 0, 0}, {B0, B0, 0, 0}, 0, 128, 1, "halo pattern + commit every 4 steps"},
        {{128, 64, 0, 0}, {0, 64, 0, 0}, {A0 + 384, A1 + 384, 0, 0}, {B0, B0, 0, 0}, 0, 128, 1, "halo pattern, shifted + commit every 4 steps"},
        {{64, 64, 64, 0}, {0, 0, 0, 0}, {A0 + 384, A0 + 384, A1 + 384, 0}, {B0 + 640, B1 + 640, B0 + 640, 0}, 1, 128, 0, "MN-major 3 x N=64, both operands row-shifted"},
    };
    long long* out;
    cudaMalloc(&out, 8);
    const int smem = 160 * 1024 + 64 + 1024;
    const int steps = 4096;
    for (const Cfg& c : cfgs) {
        int nm = 0; double math = 0;
        for (int j = 0; j < 4; ++j) if (c.n[j]) { ++nm; math += 128.0 * c.n[j] / 256.0; }
        if (nm == 1) { cudaFuncSetAttribute(umma_bench_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); umma_bench_kernel<1><<<148, 128, smem>>>(c, steps, out); }
        if (nm == 2) { cudaFuncSetAttribute(umma_bench_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); umma_bench_kernel<2><<<148, 128, smem>>>(c, steps, out); }
        if (nm == 3) { cudaFuncSetAttribute(umma_bench_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); umma_bench_kernel<3><<<148, 128, smem>>>(c, steps, out); }
        if (nm == 4) { cudaFuncSetAttribute(umma_bench_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); umma_bench_kernel<4><<<148, 128, smem>>>(c, steps, out); }
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("%s: %s\n", c.name, cudaGetErrorString(e)); return 1; }
        long long cyc;
        cudaMemcpy(&cyc, out, 8, cudaMemcpyDeviceToHost);
        printf("%-62s %7.1f cyc/step  (%d MMAs, math floor %5.0f)  -> %5.1f %% of floor rate\n", c.name, (double)cyc / steps, nm, math,
               100.0 * math / ((double)cyc / steps));
    }
    return 0;
}

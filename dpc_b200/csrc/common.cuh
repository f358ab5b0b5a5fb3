// Shared helpers for the dpc_b200 C-ABI library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/dpc_b200.h"

#define DPC_OK 0
#define DPC_ERR_ARG 1
#define DPC_ERR_CUDA 2
#define DPC_ERR_UNSUPPORTED 3

void dpc_set_error(const char* fmt, ...);

#define DPC_REQUIRE(cond, ...)                                   \
    do {                                                         \
        if (!(cond)) {                                           \
            dpc_set_error(__VA_ARGS__);                          \
            return DPC_ERR_ARG;                                  \
        }                                                        \
    } while (0)

#define DPC_CUDA(call)                                                                     \
    do {                                                                                   \
        cudaError_t e__ = (call);                                                          \
        if (e__ != cudaSuccess) {                                                          \
            dpc_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
            return DPC_ERR_CUDA;                                                           \
        }                                                                                  \
    } while (0)

void dpc_count_launch(int n);
#define DPC_LAUNCH_CHECK()          \
    do {                            \
        dpc_count_launch(1);        \
        DPC_CUDA(cudaGetLastError()); \
    } while (0)

static inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

static inline int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

int dpc_num_sms();

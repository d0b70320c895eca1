// common.cuh -- shared helpers for libsaunet_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <stdlib.h>
#include "../../include/saunet_b200.h"

namespace saunet {

void set_error(const char* fmt, ...);
void count_launch(int n = 1);
void note_kernel(const char* name);          // name of the kernel family the last C-ABI call launched (saunet_last_kernel)

#define SAUNET_CHECK_ARG(cond, code, ...)                         \
    do {                                                          \
        if (!(cond)) { saunet::set_error(__VA_ARGS__); return (code); } \
    } while (0)

#define SAUNET_CHECK_LAUNCH(name)                                                      \
    do {                                                                               \
        cudaError_t e__ = cudaGetLastError();                                          \
        if (e__ != cudaSuccess) {                                                      \
            saunet::set_error("%s: CUDA launch failed: %s", name, cudaGetErrorString(e__)); \
            return SAUNET_ERR_CUDA;                                                    \
        }                                                                              \
        saunet::count_launch();                                                        \
        saunet::note_kernel(name);                                                     \
    } while (0)

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
static inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }
static constexpr int kNumSMs = 148;
// debugging switches (environment variables) are read ONCE per process, never on the launch path
#define SAUNET_ENV_FLAG(var) ([]() -> bool { static const bool v = getenv(var) != nullptr; return v; }())

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float sigmoidf_(float v) { return 1.0f / (1.0f + __expf(-v)); }
// accurate sigmoid (matches torch.sigmoid to ~1 ulp): used where parity matters
__device__ __forceinline__ float sigmoid_acc(float v) { return 1.0f / (1.0f + expf(-v)); }

__device__ __forceinline__ float apply_act(float v, int act) {
    if (act == SAUNET_ACT_RELU) return fmaxf(v, 0.0f);
    if (act == SAUNET_ACT_SIGMOID) return sigmoid_acc(v);
    return v;
}

}  // namespace saunet

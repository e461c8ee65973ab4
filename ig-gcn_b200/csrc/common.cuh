// Shared helpers for the igcn_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/igcn_b200.h"

namespace igcn {

void set_error(const char* fmt, ...);
int sm_count();

#define IGCN_REQUIRE(cond, code, ...)          \
    do {                                       \
        if (!(cond)) {                         \
            igcn::set_error(__VA_ARGS__);      \
            return (code);                     \
        }                                      \
    } while (0)

#define IGCN_CHECK_LAUNCH(name)                                                        \
    do {                                                                               \
        cudaError_t e_ = cudaGetLastError();                                           \
        if (e_ != cudaSuccess) {                                                       \
            igcn::set_error("%s: launch failed: %s", (name), cudaGetErrorString(e_));  \
            return IGCN_ERR_LAUNCH;                                                    \
        }                                                                              \
    } while (0)

// Opt a kernel into > 48 KB of dynamic shared memory (idempotent, cheap).
template <typename K>
inline int allow_smem(K kernel, size_t bytes, const char* name) {
    if (bytes > 227 * 1024) {
        set_error("%s: needs %zu B of shared memory per CTA (> 227 KB)", name, bytes);
        return IGCN_ERR_UNSUPPORTED;
    }
    if (bytes > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
        if (e != cudaSuccess) {
            set_error("%s: cudaFuncSetAttribute(%zu B): %s", name, bytes, cudaGetErrorString(e));
            return IGCN_ERR_UNSUPPORTED;
        }
    }
    return IGCN_OK;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ float sigmoidf_(float z) { return 1.0f / (1.0f + __expf(-z)); }

}  // namespace igcn

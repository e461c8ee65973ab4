// Shared helpers for the igcn_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/igcn_b200.h"

namespace igcn {

void set_error(const char* fmt, ...);
int sm_count();
long long sm_clock_khz();

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

// ---- programmatic dependent launch ------------------------------------------------------------------------------------------------
// The benchmarked step is ~130 kernels of 2-30 us on a few dependent chains: the launch latency between two dependent kernels is
// paid ~60 times per step.  Every kernel starts with IGCN_PDL_SYNC(): it first lets the NEXT kernel of the stream be scheduled
// (griddepcontrol.launch_dependents: that grid's CTAs become resident and stop in their own IGCN_PDL_SYNC) and then waits until every
// kernel it depends on has completed and flushed its memory (griddepcontrol.wait).  Both instructions are no-ops for a kernel that
// was launched without the attribute.  launch_k() adds cudaLaunchAttributeProgrammaticStreamSerialization when IGCN_PDL=1; inside a
// stream capture the edge becomes a programmatic dependency of the CUDA graph.  Semantics are unchanged: no kernel touches global
// memory before its wait returns.
#define IGCN_PDL_SYNC()                                               \
    do {                                                              \
        asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); \
        asm volatile("griddepcontrol.wait;" ::: "memory");            \
    } while (0)

bool pdl_enabled();

template <typename... Params, typename... Args>
inline void launch_k(void (*kernel)(Params...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
    if (!pdl_enabled()) {
        kernel<<<grid, block, smem, st>>>(static_cast<Params>(args)...);
        return;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaLaunchKernelEx(&cfg, kernel, static_cast<Params>(args)...);
}

// CTAs for a persistent kernel that walks `units` work items with at most `slots` CTAs resident: the same number of rounds as
// min(slots, units) CTAs would need, but no more CTAs than that many rounds require -- 512 graphs on 444 slots are two rounds either
// way, and 256 CTAs leave 40 % of the SM slots to the kernels of the other streams (the device timeline of the config-2 step showed
// the GO backward chain waiting for an SGCN backward that had filled every SM: profiles/r2_timeline_config2_*.json).
static inline int64_t balanced_ctas(int64_t slots, int64_t units) {
    if (slots < 1) slots = 1;
    if (units <= slots) return units < 1 ? 1 : units;
    const int64_t rounds = (units + slots - 1) / slots;
    return (units + rounds - 1) / rounds;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ float sigmoidf_(float z) { return 1.0f / (1.0f + __expf(-z)); }

// grads[j] = sum over rows of partials[row][j], in a FIXED order (deterministic, no float atomics):
// a block owns 32 consecutive columns; warp w adds rows w, w+8, ... (coalesced 128 B reads), then the 8 warp
// sums are added in warp order.  The accumulation runs in fp64: with hundreds to thousands of per-CTA partials of mixed sign
// (one per subject in the GO layers) an fp32 running sum loses ~1e-4 of a gradient to cancellation (measured at B=256 against
// the fp64 oracle); the kernel reads n_rows * P floats once, so the fp64 adds are free.  Launch: <<<ceil(P/32), 256>>>.
static inline int reduce_threads(int64_t n_rows) { return n_rows > 64 ? 1024 : 256; }
static __global__ void __launch_bounds__(1024) reduce_partials_kernel(const float* __restrict__ partials, int n_rows, int P,
                                                                      float* __restrict__ grads) {
    IGCN_PDL_SYNC();
    __shared__ double sm[32][33];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const int j = blockIdx.x * 32 + lane;
    double s = 0.0;
    if (j < P) {
        const float* pp = partials + j;
        int r = warp;
        for (; r + 3 * nw < n_rows; r += 4 * nw) {            // four rows per warp in flight
            const float v0 = pp[(int64_t)r * P], v1 = pp[(int64_t)(r + nw) * P], v2 = pp[(int64_t)(r + 2 * nw) * P], v3 = pp[(int64_t)(r + 3 * nw) * P];
            s += (double)v0;
            s += (double)v1;
            s += (double)v2;
            s += (double)v3;
        }
        for (; r < n_rows; r += nw) s += (double)pp[(int64_t)r * P];
    }
    sm[warp][lane] = s;
    __syncthreads();
    if (warp == 0 && j < P) {
        double t = 0.0;
        for (int w = 0; w < nw; ++w) t += sm[w][lane];
        grads[j] = (float)t;
    }
}

}  // namespace igcn

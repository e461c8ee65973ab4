// All dropout masks of one forward pass in ONE launch (reference: nn.Dropout2d(0.4) x4 and nn.Dropout(0.5) x3 inside
// kernel/go_model.py:104,113,127,135,142 plus F.dropout p=0.5 / 0.3 at kernel/sgcn_img_snp.py:290,300 -- nine masks per pass;
// drawn with torch that is ~3 launches per mask).  Masks are multiplicative scale tensors: 0 or 1/keep.
// Philox4x32-10 (curand device API): subsequence = thread id, offset = 4 x a device-resident call counter (the offset counts
// single 32-bit outputs and every thread consumes four per call, so consecutive calls use disjoint counter blocks), so the
// launch is replayable inside a captured CUDA graph and still draws fresh, uncorrelated numbers on every replay.
#include <curand_kernel.h>

#include "common.cuh"

namespace igcn {

constexpr int kMaxSeg = 32;
struct MaskPlan {
    int64_t end[kMaxSeg];   // exclusive prefix ends (elements) of every segment inside the flat output
    float keep[kMaxSeg];    // keep probability of the segment
    int nseg;
};

__global__ void __launch_bounds__(256) dropout_masks_kernel(float* __restrict__ out, MaskPlan plan, int64_t total, uint64_t seed,
                                                            const unsigned long long* __restrict__ counter) {
    IGCN_PDL_SYNC();
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t i0 = t * 4;
    if (i0 >= total) return;
    curandStatePhilox4_32_10_t st;
    curand_init((unsigned long long)seed, (unsigned long long)t, counter[0] * 4ull, &st);
    const float4 u = curand_uniform4(&st);          // (0, 1]
    const float uv[4] = {u.x, u.y, u.z, u.w};
    int seg = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int64_t i = i0 + j;
        if (i >= total) break;
        while (seg < plan.nseg - 1 && i >= plan.end[seg]) ++seg;
        const float keep = plan.keep[seg];
        out[i] = (uv[j] <= keep) ? 1.f / keep : 0.f;
    }
}

__global__ void counter_inc_kernel(unsigned long long* counter) {
    IGCN_PDL_SYNC(); counter[0] += 1ull; }

}  // namespace igcn

extern "C" int igcn_dropout_masks(float* out, const int64_t* host_seg_end, const float* host_seg_keep, int64_t nseg, uint64_t seed,
                                  unsigned long long* counter, void* stream) {
    using namespace igcn;
    IGCN_REQUIRE(nseg > 0 && nseg <= kMaxSeg, IGCN_ERR_UNSUPPORTED, "dropout_masks: 1..%d segments supported", kMaxSeg);
    IGCN_REQUIRE(out && host_seg_end && host_seg_keep && counter, IGCN_ERR_BAD_ARG, "dropout_masks: null pointer");
    MaskPlan plan;
    plan.nseg = (int)nseg;
    int64_t prev = 0;
    for (int i = 0; i < nseg; ++i) {
        IGCN_REQUIRE(host_seg_end[i] >= prev && host_seg_keep[i] > 0.f && host_seg_keep[i] <= 1.f, IGCN_ERR_BAD_ARG,
                     "dropout_masks: segment ends must be non-decreasing and keep in (0,1]");
        plan.end[i] = prev = host_seg_end[i];
        plan.keep[i] = host_seg_keep[i];
    }
    const int64_t total = prev;
    cudaStream_t st = (cudaStream_t)stream;
    if (total > 0) {
        const int64_t threads = (total + 3) / 4;
        igcn::launch_k(dropout_masks_kernel, dim3((unsigned)((threads + 255) / 256)), dim3(256), 0, st, out, plan, total, seed, counter);
        IGCN_CHECK_LAUNCH("dropout_masks");
    }
    igcn::launch_k(counter_inc_kernel, dim3(1), dim3(1), 0, st, counter);
    IGCN_CHECK_LAUNCH("dropout_counter_inc");
    return IGCN_OK;
}

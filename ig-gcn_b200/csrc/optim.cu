// Fused flat-buffer Adam: ONE launch updates every parameter of the model (reference: torch.optim.Adam as used at
// kernel/train_eval_sgcn_img_snps.py:108,547 -- lr 1e-3, betas (0.9,0.999), eps 1e-8, weight_decay 0, no amsgrad).
// Parameters, gradients and both moments are contiguous fp32 buffers (the gradient buffer is the one the data-parallel
// all-reduce runs on).  `step` and `lr` live in device memory so the launch can sit inside a captured CUDA graph.
#include "common.cuh"

namespace igcn {

__global__ void __launch_bounds__(256) adam_flat_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                        float* __restrict__ v, const float* __restrict__ step,
                                                        const float* __restrict__ lr, float beta1, float beta2, float eps,
                                                        float grad_scale, int64_t n) {
    const float t = step[0];                       // already incremented for this update (1, 2, ...)
    const float bias1 = 1.f - powf(beta1, t);
    const float bias2 = 1.f - powf(beta2, t);
    const float step_size = lr[0] / bias1;
    const float inv_sqrt_bias2 = rsqrtf(bias2);
    const int64_t n4 = n >> 2;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        float4 pv = reinterpret_cast<float4*>(p)[i];
        const float4 gv = reinterpret_cast<const float4*>(g)[i];
        float4 mv = reinterpret_cast<float4*>(m)[i];
        float4 vv = reinterpret_cast<float4*>(v)[i];
#define IGCN_ADAM1(C)                                                   \
    {                                                                   \
        const float gg = gv.C * grad_scale;                             \
        mv.C = beta1 * mv.C + (1.f - beta1) * gg;                       \
        vv.C = beta2 * vv.C + (1.f - beta2) * gg * gg;                  \
        pv.C -= step_size * mv.C / (sqrtf(vv.C) * inv_sqrt_bias2 + eps); \
    }
        IGCN_ADAM1(x) IGCN_ADAM1(y) IGCN_ADAM1(z) IGCN_ADAM1(w)
        reinterpret_cast<float4*>(p)[i] = pv;
        reinterpret_cast<float4*>(m)[i] = mv;
        reinterpret_cast<float4*>(v)[i] = vv;
    }
    for (int64_t i = (n4 << 2) + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const float gg = g[i] * grad_scale;
        const float mm = beta1 * m[i] + (1.f - beta1) * gg;
        const float vv = beta2 * v[i] + (1.f - beta2) * gg * gg;
        m[i] = mm;
        v[i] = vv;
        p[i] -= step_size * mm / (sqrtf(vv) * inv_sqrt_bias2 + eps);
    }
#undef IGCN_ADAM1
}

// Gradient gather: copies up to GATHER_MAX separate gradient tensors into their slots of the flat buffer in ONE launch
// (a null source zero-fills the slot).  The table travels by value, so the launch can sit in a captured graph.
constexpr int GATHER_MAX = 96;
struct GatherTable {
    const float* src[GATHER_MAX];
    int64_t off[GATHER_MAX];
    int n[GATHER_MAX];
};

__global__ void __launch_bounds__(256) gather_flat_kernel(const __grid_constant__ GatherTable t, float* __restrict__ dst,
                                                          float* __restrict__ step_counter) {
    if (step_counter && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) step_counter[0] += 1.f;
    const int e = blockIdx.y;
    const float* __restrict__ s = t.src[e];
    float* __restrict__ d = dst + t.off[e];
    const int n = t.n[e];
    const int stride = gridDim.x * 256;
    const int i0 = blockIdx.x * 256 + threadIdx.x;
    if (s != nullptr && (reinterpret_cast<uintptr_t>(s) & 15) == 0) {
        const int n4 = n >> 2;
        for (int i = i0; i < n4; i += 4 * stride) {                 // 4 x 16-byte loads in flight per thread
            float4 v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (i + u * stride < n4) v[u] = reinterpret_cast<const float4*>(s)[i + u * stride];
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (i + u * stride < n4) reinterpret_cast<float4*>(d)[i + u * stride] = v[u];
        }
        for (int i = (n4 << 2) + i0; i < n; i += stride) d[i] = s[i];
    } else {
        for (int i = i0; i < n; i += stride) d[i] = s ? s[i] : 0.f;
    }
}

}  // namespace igcn

extern "C" int igcn_gather_flat(const int64_t* host_src_ptrs, const int64_t* host_offsets, const int64_t* host_sizes, int64_t count,
                                float* flat, int64_t flat_n, float* step_counter, void* stream) {
    using namespace igcn;
    IGCN_REQUIRE(count >= 0, IGCN_ERR_BAD_ARG, "gather_flat: negative count");
    if (count == 0) return IGCN_OK;
    IGCN_REQUIRE(host_src_ptrs && host_offsets && host_sizes && flat, IGCN_ERR_BAD_ARG, "gather_flat: null pointer");
    IGCN_REQUIRE((uintptr_t)flat % 16 == 0, IGCN_ERR_BAD_ARG, "gather_flat: flat buffer must be 16-byte aligned");
    for (int64_t i = 0; i < count; ++i) {
        IGCN_REQUIRE(host_sizes[i] >= 0 && host_sizes[i] < (int64_t(1) << 31) && host_offsets[i] >= 0 && (host_offsets[i] & 3) == 0 &&
                         host_offsets[i] + host_sizes[i] <= flat_n,
                     IGCN_ERR_BAD_ARG, "gather_flat: slot %lld (offset %lld, size %lld) outside the flat buffer of %lld or not 16-byte aligned",
                     (long long)i, (long long)host_offsets[i], (long long)host_sizes[i], (long long)flat_n);
        IGCN_REQUIRE(host_src_ptrs[i] % 4 == 0, IGCN_ERR_BAD_ARG, "gather_flat: source %lld is not 4-byte aligned", (long long)i);
    }
    for (int64_t base = 0; base < count; base += GATHER_MAX) {
        GatherTable t;
        const int m = (int)((count - base) < GATHER_MAX ? (count - base) : GATHER_MAX);
        int64_t biggest = 1;
        for (int e = 0; e < m; ++e) {
            t.src[e] = reinterpret_cast<const float*>(host_src_ptrs[base + e]);
            t.off[e] = host_offsets[base + e];
            t.n[e] = (int)host_sizes[base + e];
            if (host_sizes[base + e] > biggest) biggest = host_sizes[base + e];
        }
        for (int e = m; e < GATHER_MAX; ++e) { t.src[e] = nullptr; t.off[e] = 0; t.n[e] = 0; }
        int64_t gx = (biggest / 4 + 256 * 4 - 1) / (256 * 4);       // one round of 4 float4 per thread covers the largest tensor ...
        if (gx > 16) gx = 16;                                       // ... up to 16 CTAs per tensor
        if (gx < 1) gx = 1;
        gather_flat_kernel<<<dim3((unsigned)gx, (unsigned)m), 256, 0, (cudaStream_t)stream>>>(t, flat, base == 0 ? step_counter : nullptr);
        IGCN_CHECK_LAUNCH("gather_flat");
    }
    return IGCN_OK;
}

extern "C" int64_t igcn_gather_flat_launches(int64_t count) { return count <= 0 ? 0 : (count + igcn::GATHER_MAX - 1) / igcn::GATHER_MAX; }

extern "C" int igcn_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, const float* step,
                              const float* lr, double beta1, double beta2, double eps, double grad_scale, int64_t n, void* stream) {
    using namespace igcn;
    IGCN_REQUIRE(n >= 0, IGCN_ERR_BAD_ARG, "adam_step: negative size");
    if (n == 0) return IGCN_OK;
    IGCN_REQUIRE(params && grads && exp_avg && exp_avg_sq && step && lr, IGCN_ERR_BAD_ARG, "adam_step: null pointer");
    IGCN_REQUIRE(((uintptr_t)params | (uintptr_t)grads | (uintptr_t)exp_avg | (uintptr_t)exp_avg_sq) % 16 == 0, IGCN_ERR_BAD_ARG,
                 "adam_step: buffers must be 16-byte aligned");
    int64_t blocks = (n / 4 + 255) / 256;
    const int64_t cap = (int64_t)sm_count() * 8;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    adam_flat_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(params, grads, exp_avg, exp_avg_sq, step, lr, (float)beta1,
                                                                    (float)beta2, (float)eps, (float)grad_scale, n);
    IGCN_CHECK_LAUNCH("adam_step");
    return IGCN_OK;
}

// =============================================================================================================================
// Data-parallel step: gradient all-reduce + Adam as ONE kernel over NVLink peer memory.
//
// Every rank keeps its flat gradient buffer in symmetric (peer-mapped) memory.  The kernel of rank r
//   1. signals every peer and waits for every peer, block by block (all gradient buffers are complete and visible),
//   2. for the slice it OWNS of every block's range: loads the chunks from ALL ranks (its own through L2, the others over
//      NVLink / NVSwitch), adds them in rank order and stores the sum back into all ranks' buffers, in place -- every sum is computed
//      once, so the replicas are bit-identical and never drift, and a rank moves n bytes each way whatever the world size,
//   3. signals / waits again (all sums have landed everywhere; from here on a rank touches only its own memory),
//   4. scales its now-reduced gradient by 1/world and applies the Adam update to its own parameter replica.
// This replaces ncclAllReduce + a separate optimizer launch (at 1.66 MB per rank the collective is latency bound: two device-side
// flag exchanges and one pass over the data instead of a ring).  Flags are one 32-bit word per (block, peer) in each rank's signal
// pad, set with a release CAS 0 -> 1 by the sender and cleared with an acquire CAS 1 -> 0 by the receiver, so they reset themselves
// and the launch can be replayed inside a CUDA graph.  Waits are bounded by a caller-chosen timeout (minutes by default: ranks may
// legitimately be seconds apart -- logging, evaluation, lazy module loads); on expiry the kernel records the peer in a
// host-checked error flag and leaves the wait (the step's result is then invalid but the CUDA context survives and the host
// can raise); without a flag it traps.  All ranks must issue their steps in lock-step, one fused launch per rank per step.
// =============================================================================================================================
namespace igcn {

constexpr int DP_MAX_WORLD = 16;
struct DpPeers {
    float* grad[DP_MAX_WORLD];           // gradient buffer of every rank (peer pointers)
    uint32_t* signal[DP_MAX_WORLD];      // signal pad of every rank
};

__device__ __forceinline__ uint32_t cas_release_sys(uint32_t* addr, uint32_t cmp, uint32_t val) {
    uint32_t old;
    asm volatile("atom.global.release.sys.cas.b32 %0, [%1], %2, %3;" : "=r"(old) : "l"(addr), "r"(cmp), "r"(val) : "memory");
    return old;
}
__device__ __forceinline__ uint32_t cas_acquire_sys(uint32_t* addr, uint32_t cmp, uint32_t val) {
    uint32_t old;
    asm volatile("atom.global.acquire.sys.cas.b32 %0, [%1], %2, %3;" : "=r"(old) : "l"(addr), "r"(cmp), "r"(val) : "memory");
    return old;
}

// all blocks with the same blockIdx.x on all ranks meet here; slot = which of the kernel's barriers (distinct flag words)
__device__ __forceinline__ void dp_timeout(int* error_flag, int peer, int what) {
    if (error_flag) {
        atomicExch(error_flag, (what << 8) | (peer + 1));           // host side: FlatAdam.check_dp_error()
        __threadfence_system();
    } else {
        printf("igcn dp_adam: %s rank %d timed out\n", what == 1 ? "signal to" : "wait for", peer);
        __trap();
    }
}

__device__ __forceinline__ void dp_block_barrier(const DpPeers& peers, int rank, int world, int slot, long long timeout_cycles,
                                                 int* error_flag) {
    __syncthreads();
    if ((int)threadIdx.x < world) {
        const int peer = threadIdx.x;
        const size_t base = ((size_t)slot * gridDim.x + blockIdx.x) * world;
        uint32_t* put = peers.signal[peer] + base + rank;          // my flag in the peer's pad
        uint32_t* get = peers.signal[rank] + base + peer;          // the peer's flag in my pad
        long long t0 = clock64();
        while (cas_release_sys(put, 0u, 1u) != 0u)
            if (clock64() - t0 > timeout_cycles) { dp_timeout(error_flag, peer, 1); break; }
        t0 = clock64();                                             // each phase gets the full budget
        while (cas_acquire_sys(get, 1u, 0u) != 1u)
            if (clock64() - t0 > timeout_cycles) { dp_timeout(error_flag, peer, 2); break; }
    }
    __syncthreads();
}

__global__ void __launch_bounds__(512) dp_allreduce_adam_kernel(DpPeers peers, int rank, int world, float* __restrict__ p,
                                                                float* __restrict__ m, float* __restrict__ v,
                                                                const float* __restrict__ step, const float* __restrict__ lr, float beta1,
                                                                float beta2, float eps, int64_t n, long long timeout_cycles,
                                                                int* error_flag) {
    dp_block_barrier(peers, rank, world, 0, timeout_cycles, error_flag);
    const float t = step[0];
    const float bias1 = 1.f - powf(beta1, t), bias2 = 1.f - powf(beta2, t);
    const float step_size = lr[0] / bias1, inv_sqrt_bias2 = rsqrtf(bias2), scale = 1.f / (float)world;
    // block b of every rank works on the same contiguous range of 16-byte chunks [c0, c1); n is a multiple of 4 (FlatAdam pads)
    const int64_t n4 = n >> 2, per = (n4 + gridDim.x - 1) / gridDim.x;
    const int64_t c0 = min(n4, (int64_t)blockIdx.x * per), c1 = min(n4, c0 + per);
    float* __restrict__ mine = peers.grad[rank];
    if (world > 1) {
        // Phase 1 (reduce-scatter + all-gather in place): rank r OWNS the r-th slice of the block's range; it reads the slice from all
        // ranks, adds in rank order and stores the sum back into ALL ranks' buffers.  NVLink traffic per rank: n read + n written,
        // independent of the world size (the one-shot version of round 1 read world * n per rank), every sum is computed once, and
        // all threads of the block share the slice, so the phase is one or two round trips deep.
        const int64_t slice = (c1 - c0 + world - 1) / world;
        const int64_t s0 = min(c1, c0 + (int64_t)rank * slice), s1 = min(c1, s0 + slice);
        for (int64_t i = s0 + threadIdx.x; i < s1; i += blockDim.x) {
            float4 x[DP_MAX_WORLD];
#pragma unroll
            for (int r = 0; r < DP_MAX_WORLD; ++r)
                if (r < world) x[r] = reinterpret_cast<const float4*>(peers.grad[r])[i];
            float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int r = 0; r < DP_MAX_WORLD; ++r)                              // rank order: the same sum as a serial loop over ranks
                if (r < world) { g.x += x[r].x; g.y += x[r].y; g.z += x[r].z; g.w += x[r].w; }
#pragma unroll
            for (int r = 0; r < DP_MAX_WORLD; ++r)
                if (r < world) reinterpret_cast<float4*>(peers.grad[r])[i] = g;
        }
        // every rank's sums of this block's chunks have landed everywhere; after this point a rank touches only its own memory, so
        // no closing barrier is needed (the next step's opening barrier orders the next gather against these reads)
        dp_block_barrier(peers, rank, world, 1, timeout_cycles, error_flag);
    }
    for (int64_t i = c0 + threadIdx.x; i < c1; i += blockDim.x) {
        const float4 g = __ldcg(reinterpret_cast<const float4*>(mine) + i);      // written by the owner rank: not through L1
        float4 pv = reinterpret_cast<float4*>(p)[i], mv = reinterpret_cast<float4*>(m)[i], vv = reinterpret_cast<float4*>(v)[i];
#define IGCN_ADAM1(C)                                                   \
    {                                                                   \
        const float gg = g.C * scale;                                   \
        mv.C = beta1 * mv.C + (1.f - beta1) * gg;                       \
        vv.C = beta2 * vv.C + (1.f - beta2) * gg * gg;                  \
        pv.C -= step_size * mv.C / (sqrtf(vv.C) * inv_sqrt_bias2 + eps); \
    }
        IGCN_ADAM1(x) IGCN_ADAM1(y) IGCN_ADAM1(z) IGCN_ADAM1(w)
#undef IGCN_ADAM1
        reinterpret_cast<float4*>(p)[i] = pv;
        reinterpret_cast<float4*>(m)[i] = mv;
        reinterpret_cast<float4*>(v)[i] = vv;
    }
}

}  // namespace igcn

extern "C" int64_t igcn_dp_adam_blocks(int64_t n, int64_t world, int64_t signal_pad_bytes) {
    using namespace igcn;
    if (world < 1) world = 1;
    int64_t blocks = (n / 4 + 511) / 512;
    const int64_t by_pad = signal_pad_bytes / (2 * 4 * world);                 // two barriers x 4 bytes x world flags per block
    if (blocks > by_pad) blocks = by_pad;
    if (blocks > sm_count()) blocks = sm_count();                              // all blocks of all ranks must be co-resident
    return blocks < 1 ? 0 : blocks;
}

extern "C" int igcn_dp_allreduce_adam(const int64_t* host_grad_ptrs, const int64_t* host_signal_ptrs, int64_t rank, int64_t world,
                                      int64_t signal_pad_bytes, float* params, float* exp_avg, float* exp_avg_sq, const float* step,
                                      const float* lr, double beta1, double beta2, double eps, int64_t n, int64_t timeout_ms,
                                      int* error_flag, void* stream) {
    using namespace igcn;
    IGCN_REQUIRE(timeout_ms > 0, IGCN_ERR_BAD_ARG, "dp_allreduce_adam: timeout_ms must be positive");
    IGCN_REQUIRE(host_grad_ptrs && host_signal_ptrs && params && exp_avg && exp_avg_sq && step && lr, IGCN_ERR_BAD_ARG, "dp_allreduce_adam: null pointer");
    IGCN_REQUIRE(world >= 1 && world <= DP_MAX_WORLD && rank >= 0 && rank < world, IGCN_ERR_BAD_ARG, "dp_allreduce_adam: rank %lld of %lld",
                 (long long)rank, (long long)world);
    IGCN_REQUIRE(n > 0 && (n & 3) == 0, IGCN_ERR_BAD_ARG, "dp_allreduce_adam: n must be a positive multiple of 4");
    const int64_t blocks = igcn_dp_adam_blocks(n, world, signal_pad_bytes);
    IGCN_REQUIRE(blocks >= 1, IGCN_ERR_UNSUPPORTED, "dp_allreduce_adam: signal pad of %lld bytes is too small for %lld ranks",
                 (long long)signal_pad_bytes, (long long)world);
    DpPeers peers;
    for (int r = 0; r < DP_MAX_WORLD; ++r) {
        peers.grad[r] = r < world ? reinterpret_cast<float*>(host_grad_ptrs[r]) : nullptr;
        peers.signal[r] = r < world ? reinterpret_cast<uint32_t*>(host_signal_ptrs[r]) : nullptr;
        IGCN_REQUIRE(r >= world || (peers.grad[r] && peers.signal[r] && ((uintptr_t)peers.grad[r] & 15) == 0), IGCN_ERR_BAD_ARG,
                     "dp_allreduce_adam: bad peer pointer for rank %d", r);
    }
    dp_allreduce_adam_kernel<<<(unsigned)blocks, 512, 0, (cudaStream_t)stream>>>(peers, (int)rank, (int)world, params, exp_avg, exp_avg_sq, step, lr,
                                                                                (float)beta1, (float)beta2, (float)eps, n,
                                                                                (long long)timeout_ms * sm_clock_khz(), error_flag);
    IGCN_CHECK_LAUNCH("dp_allreduce_adam");
    return IGCN_OK;
}

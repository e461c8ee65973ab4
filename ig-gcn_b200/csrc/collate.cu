// Device-side batched-graph collation: COO -> (target-sorted CSR, source-sorted transposed index).
// Replaces Batch.from_data_list (reference batch.py:24-123) and the index work PyG's gcn_norm /
// propagate redo on every GCNConv call.  Integer work: results are bit exact with the reference
// (edge_index, batch) and with the stable-sort layout contract (oracle: target_sorted_csr).
//
// One CTA per graph (graphs own R consecutive node ids and a contiguous edge range), counting sort in
// shared memory.  Slot assignment uses shared-memory INTEGER atomics, then each row is put into original
// edge order by its owning thread, so the result is deterministic and equals a stable sort.
#include "common.cuh"

namespace igcn {

template <typename IdxT>
struct EdgeIn {
    const IdxT* src;     // LOCAL ids when IdxT=int32_t; GLOBAL ids when int64_t
    const IdxT* dst;
};

// shared layout (ints): cnt_t[R] | cnt_s[R] | start_t[R+1] | start_s[R+1] | key_t | key_s | tmp_t | tmp_s  (maxEg each)
__device__ __forceinline__ void block_exclusive_scan(const int* cnt, int* start, int n) {
    // n is small (<= a few thousand): chunked warp scan executed by ONE (converged) warp.
    const int lane = threadIdx.x & 31;
    {
        int carry = 0;
        for (int base = 0; base < n; base += 32) {
            int i = base + lane;
            int v = (i < n) ? cnt[i] : 0;
            int inc = v;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                int t = __shfl_up_sync(0xffffffffu, inc, o);
                if (lane >= o) inc += t;
            }
            if (i < n) start[i] = carry + inc - v;
            carry += __shfl_sync(0xffffffffu, inc, 31);
        }
        if (lane == 0) start[n] = carry;
    }
}

__device__ __forceinline__ void sort_row(int* a, int n) {  // ascending insertion sort; rows are short (top-k in-edges)
    for (int i = 1; i < n; ++i) {
        int v = a[i], j = i - 1;
        while (j >= 0 && a[j] > v) {
            a[j + 1] = a[j];
            --j;
        }
        a[j + 1] = v;
    }
}

template <typename IdxT, bool kGlobalIn, bool kWriteBatch>
__global__ void __launch_bounds__(256) collate_kernel(EdgeIn<IdxT> in, const float* __restrict__ w,
                                                      const int64_t* __restrict__ gptr64, const int32_t* __restrict__ gptr32,
                                                      int B, int R, int64_t E, int maxEg,
                                                      int64_t* __restrict__ edge_index, int64_t* __restrict__ batch,
                                                      int32_t* __restrict__ rowptr_t, int32_t* __restrict__ csr_src,
                                                      int32_t* __restrict__ csr_perm, float* __restrict__ csr_w,
                                                      int32_t* __restrict__ rowptr_s, int32_t* __restrict__ csc_pos) {
    extern __shared__ int sm[];
    int* cnt_t = sm;
    int* cnt_s = cnt_t + R;
    int* start_t = cnt_s + R;
    int* start_s = start_t + R + 1;
    int* key_t = start_s + R + 1;   // target of edge k (later reused as slot_of[k])
    int* key_s = key_t + maxEg;     // source of edge k
    int* tmp_t = key_s + maxEg;     // edge ids grouped by target
    int* tmp_s = tmp_t + maxEg;     // edge ids grouped by source
    const int tid = threadIdx.x, nt = blockDim.x;

    for (int g = blockIdx.x; g < B; g += gridDim.x) {
        const int64_t e0 = gptr64 ? gptr64[g] : (int64_t)gptr32[g];
        const int64_t e1 = gptr64 ? gptr64[g + 1] : (int64_t)gptr32[g + 1];
        const int Eg = (int)(e1 - e0);
        const int64_t node0 = (int64_t)g * R;
        if (Eg > maxEg) __trap();  // host contract violated (max_eg too small)

        for (int i = tid; i < R; i += nt) {
            cnt_t[i] = 0;
            cnt_s[i] = 0;
            if (kWriteBatch) batch[node0 + i] = g;
        }
        __syncthreads();
        for (int k = tid; k < Eg; k += nt) {
            int s, t;
            if (kGlobalIn) {
                s = (int)((int64_t)in.src[e0 + k] - node0);
                t = (int)((int64_t)in.dst[e0 + k] - node0);
            } else {
                s = (int)in.src[e0 + k];
                t = (int)in.dst[e0 + k];
                edge_index[e0 + k] = node0 + s;
                edge_index[E + e0 + k] = node0 + t;
            }
            // host contract (validated in SubjectSet / Batch.from_device_tensors): every edge stays inside its graph
            if ((unsigned)s >= (unsigned)R || (unsigned)t >= (unsigned)R) __trap();
            key_t[k] = t;
            key_s[k] = s;
            atomicAdd(&cnt_t[t], 1);   // integer atomics: the counts are order independent
            atomicAdd(&cnt_s[s], 1);
        }
        __syncthreads();
        if ((tid >> 5) == 0) block_exclusive_scan(cnt_t, start_t, R);
        if ((tid >> 5) == (nt > 32 ? 1 : 0)) block_exclusive_scan(cnt_s, start_s, R);
        __syncthreads();
        for (int i = tid; i < R; i += nt) {
            rowptr_t[node0 + i] = (int32_t)(e0 + start_t[i]);
            rowptr_s[node0 + i] = (int32_t)(e0 + start_s[i]);
            cnt_t[i] = 0;  // reuse as fill cursors
            cnt_s[i] = 0;
        }
        if (g == B - 1 && tid == 0) {
            rowptr_t[node0 + R] = (int32_t)e1;
            rowptr_s[node0 + R] = (int32_t)e1;
        }
        __syncthreads();
        // drop every edge id into its row (slot inside the row is arbitrary here ...)
        for (int k = tid; k < Eg; k += nt) {
            const int t = key_t[k], s = key_s[k];
            tmp_t[start_t[t] + atomicAdd(&cnt_t[t], 1)] = k;
            tmp_s[start_s[s] + atomicAdd(&cnt_s[s], 1)] = k;
        }
        __syncthreads();
        // ... and is made deterministic here: ascending edge id inside every row == stable sort
        for (int i = tid; i < R; i += nt) {
            sort_row(tmp_t + start_t[i], start_t[i + 1] - start_t[i]);
            sort_row(tmp_s + start_s[i], start_s[i + 1] - start_s[i]);
        }
        __syncthreads();
        int* slot_of = key_t;
        for (int q = tid; q < Eg; q += nt) {
            const int k = tmp_t[q];
            csr_src[e0 + q] = (int32_t)(node0 + key_s[k]);
            csr_perm[e0 + q] = (int32_t)(e0 + k);
            csr_w[e0 + q] = w[e0 + k];
        }
        __syncthreads();
        for (int q = tid; q < Eg; q += nt) slot_of[tmp_t[q]] = q;
        __syncthreads();
        for (int q = tid; q < Eg; q += nt) csc_pos[e0 + q] = (int32_t)(e0 + slot_of[tmp_s[q]]);
        __syncthreads();
    }
}

// first edge whose source belongs to graph >= g  (edges of a collated batch are grouped by graph)
__global__ void graph_eptr_kernel(const int64_t* __restrict__ src, int64_t E, int B, int R, int32_t* __restrict__ eptr) {
    int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g > B) return;
    int64_t lo = 0, hi = E;
    const int64_t key = (int64_t)g * R;
    while (lo < hi) {
        int64_t mid = (lo + hi) >> 1;
        if (src[mid] < key)
            lo = mid + 1;
        else
            hi = mid;
    }
    eptr[g] = (int32_t)lo;
}

static size_t collate_smem(int R, int maxEg) { return sizeof(int) * (size_t)(4 * R + 2 + 4 * maxEg); }

}  // namespace igcn

using namespace igcn;

extern "C" int igcn_collate_csr(const int64_t* graph_ptr, const int32_t* loc_src, const int32_t* loc_dst, const float* w,
                                int64_t B, int64_t R, int64_t E, int64_t max_eg, int64_t* edge_index, int64_t* batch,
                                int32_t* rowptr_t, int32_t* csr_src, int32_t* csr_perm, float* csr_w, int32_t* rowptr_s,
                                int32_t* csc_pos, void* stream) {
    IGCN_REQUIRE(B >= 0 && R > 0 && E >= 0 && max_eg >= 0, IGCN_ERR_BAD_ARG, "collate_csr: negative size");
    IGCN_REQUIRE(B * R < (1ll << 31) && E < (1ll << 31), IGCN_ERR_UNSUPPORTED, "collate_csr: >2^31 nodes or edges per batch");
    if (B == 0) return IGCN_OK;
    IGCN_REQUIRE(graph_ptr && batch && rowptr_t && rowptr_s, IGCN_ERR_BAD_ARG, "collate_csr: null pointer");
    IGCN_REQUIRE(E == 0 || edge_index, IGCN_ERR_BAD_ARG, "collate_csr: null edge_index");
    IGCN_REQUIRE(E == 0 || (loc_src && loc_dst && w && csr_src && csr_perm && csr_w && csc_pos), IGCN_ERR_BAD_ARG,
                 "collate_csr: null edge pointer");
    cudaStream_t st = (cudaStream_t)stream;
    if (B == 0) return IGCN_OK;
    auto kern = collate_kernel<int32_t, false, true>;
    size_t smem = collate_smem((int)R, (int)max_eg);
    int rc = allow_smem(kern, smem, "collate_csr");
    if (rc) return rc;
    int grid = (int)(B < (int64_t)sm_count() * 8 ? B : (int64_t)sm_count() * 8);
    EdgeIn<int32_t> in{loc_src, loc_dst};
    kern<<<grid, 256, smem, st>>>(in, w, graph_ptr, nullptr, (int)B, (int)R, E, (int)max_eg, edge_index, batch, rowptr_t,
                                  csr_src, csr_perm, csr_w, rowptr_s, csc_pos);
    IGCN_CHECK_LAUNCH("collate_csr");
    return IGCN_OK;
}

extern "C" int igcn_csr_from_edge_index(const int64_t* edge_index, const float* w, int64_t B, int64_t R, int64_t E,
                                        int64_t max_eg, int32_t* graph_eptr, int32_t* rowptr_t, int32_t* csr_src,
                                        int32_t* csr_perm, float* csr_w, int32_t* rowptr_s, int32_t* csc_pos, void* stream) {
    IGCN_REQUIRE(B >= 0 && R > 0 && E >= 0 && max_eg >= 0, IGCN_ERR_BAD_ARG, "csr_from_edge_index: negative size");
    IGCN_REQUIRE(B * R < (1ll << 31) && E < (1ll << 31), IGCN_ERR_UNSUPPORTED, "csr_from_edge_index: >2^31 nodes or edges");
    if (B == 0) return IGCN_OK;
    IGCN_REQUIRE(graph_eptr && rowptr_t && rowptr_s, IGCN_ERR_BAD_ARG, "csr_from_edge_index: null pointer");
    IGCN_REQUIRE(E == 0 || (edge_index && w && csr_src && csr_perm && csr_w && csc_pos), IGCN_ERR_BAD_ARG,
                 "csr_from_edge_index: null edge pointer");
    cudaStream_t st = (cudaStream_t)stream;
    if (B == 0) return IGCN_OK;
    graph_eptr_kernel<<<(int)((B + 1 + 127) / 128), 128, 0, st>>>(edge_index, E, (int)B, (int)R, graph_eptr);
    IGCN_CHECK_LAUNCH("graph_eptr");
    auto kern = collate_kernel<int64_t, true, false>;
    size_t smem = collate_smem((int)R, (int)max_eg);
    int rc = allow_smem(kern, smem, "csr_from_edge_index");
    if (rc) return rc;
    int grid = (int)(B < (int64_t)sm_count() * 8 ? B : (int64_t)sm_count() * 8);
    EdgeIn<int64_t> in{edge_index, edge_index + E};
    kern<<<grid, 256, smem, st>>>(in, w, nullptr, graph_eptr, (int)B, (int)R, E, (int)max_eg, nullptr, nullptr, rowptr_t,
                                  csr_src, csr_perm, csr_w, rowptr_s, csc_pos);
    IGCN_CHECK_LAUNCH("csr_from_edge_index");
    return IGCN_OK;
}

// Graph diffusion convolution, sparsification step, on the device (reference: util_gdc.py:25-31 `get_top_k_matrix` + :84-101
// dense -> COO, applied per subject by sgcn_data.py:332-338 as a CPU pre_transform).
//
// Input: the dense personalised-PageRank matrices PPR = alpha (I - (1-alpha) D^-1/2 A D^-1/2)^-1 of a batch of subjects
// (util_gdc.py:7-14; the R x R inverse itself is a batched fp64 LAPACK-style inverse -- a library call, like the reference's
// np.linalg.inv).  Per subject and per COLUMN keep the k largest entries, divide the column by their sum, cast to fp32 and emit
// the non-zeros in ROW-MAJOR order (= scipy coo_matrix(dense) order), i.e. exactly k in-edges per node:
//     edge (src = row, dst = column, weight).
// One CTA per subject; a thread owns a column for the selection (coalesced row walk), rows are then filled through a counting
// sort (integer atomics) and put in ascending column order by their owning thread, so the edge list is bit-identical to the
// numpy restatement (igcn_b200/synthetic.py::gdc_topk + make_subjects) for the same PPR matrix.
#include "common.cuh"

namespace igcn {
namespace gdc {

constexpr int KMAX = 8;

template <typename T>
__global__ void __launch_bounds__(256) topk_emit_kernel(const T* __restrict__ ppr, int B, int R, int k, int32_t* __restrict__ esrc,
                                                        int32_t* __restrict__ edst, float* __restrict__ eattr) {
    extern __shared__ int smi[];
    int* cnt = smi;                       // R      entries per row
    int* start = cnt + R;                 // R + 1
    int* cur = start + R + 1;             // R      fill cursors
    int* rcol = cur + R;                  // R * k  column of every kept entry, grouped by row
    float* rval = reinterpret_cast<float*>(rcol + R * k);     // R * k
    const int tid = threadIdx.x, nt = blockDim.x;
    for (int b = blockIdx.x; b < B; b += gridDim.x) {
        const T* P = ppr + (int64_t)b * R * R;
        for (int i = tid; i < R; i += nt) {
            cnt[i] = 0;
            cur[i] = 0;
        }
        __syncthreads();
        // selection: thread = column.  np.argsort(ppr, axis=1)[: R-k] are dropped = the k largest stay; among equal values argsort's
        // order decides in numpy -- exact ties do not occur for PPR matrices (checked by the test against the numpy restatement)
        for (int c = tid; c < R; c += nt) {
            T bv[KMAX];
            int br[KMAX];
#pragma unroll
            for (int q = 0; q < KMAX; ++q) {
                bv[q] = (T)-1;
                br[q] = -1;
            }
            for (int r = 0; r < R; ++r) {
                const T v = P[(int64_t)r * R + c];
                if (v > bv[k - 1]) {                       // insert into the sorted top-k (descending)
                    int q = k - 1;
                    while (q > 0 && v > bv[q - 1]) {
                        bv[q] = bv[q - 1];
                        br[q] = br[q - 1];
                        --q;
                    }
                    bv[q] = v;
                    br[q] = r;
                }
            }
            T sum = 0;
            for (int q = 0; q < k; ++q)
                if (br[q] >= 0) sum += bv[q];
            if (!(sum > 0)) sum = 1;                       // util_gdc.py:29 norm[norm <= 0] = 1
            for (int q = 0; q < k; ++q) {
                if (br[q] < 0) continue;
                const float w = (float)(bv[q] / sum);
                if (w != 0.f) atomicAdd(&cnt[br[q]], 1);
            }
            // keep the selection for the fill pass
            for (int q = 0; q < k; ++q) {
                rcol[c * k + q] = br[q];                   // temporarily: row of the q-th kept entry of column c
                rval[c * k + q] = br[q] >= 0 ? (float)(bv[q] / sum) : 0.f;
            }
        }
        __syncthreads();
        if (tid < 32) {                                    // exclusive scan over rows (one warp, chunked)
            int carry = 0;
            for (int base = 0; base < R; base += 32) {
                const int i = base + tid;
                const int v = i < R ? cnt[i] : 0;
                int inc = v;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int t = __shfl_up_sync(0xffffffffu, inc, o);
                    if (tid >= o) inc += t;
                }
                if (i < R) start[i] = carry + inc - v;
                carry += __shfl_sync(0xffffffffu, inc, 31);
            }
            if (tid == 0) start[R] = carry;
        }
        __syncthreads();
        // fill: every kept (row, column, weight) into its row's segment of the output (slot order arbitrary here ...)
        const int64_t e0 = (int64_t)b * R * k;
        for (int c = tid; c < R; c += nt) {
            for (int q = 0; q < k; ++q) {
                const int r = rcol[c * k + q];
                const float w = rval[c * k + q];
                if (r < 0 || w == 0.f) continue;
                const int slot = start[r] + atomicAdd(&cur[r], 1);
                esrc[e0 + slot] = r;
                edst[e0 + slot] = c;
                eattr[e0 + slot] = w;
            }
        }
        __syncthreads();
        // ... made deterministic here: ascending column inside every row = row-major COO order
        for (int r = tid; r < R; r += nt) {
            const int s0 = start[r], n = start[r + 1] - start[r];
            for (int i = 1; i < n; ++i) {
                const int c = edst[e0 + s0 + i];
                const float w = eattr[e0 + s0 + i];
                int j = i - 1;
                while (j >= 0 && edst[e0 + s0 + j] > c) {
                    edst[e0 + s0 + j + 1] = edst[e0 + s0 + j];
                    eattr[e0 + s0 + j + 1] = eattr[e0 + s0 + j];
                    --j;
                }
                edst[e0 + s0 + j + 1] = c;
                eattr[e0 + s0 + j + 1] = w;
            }
        }
        __syncthreads();
    }
}

}  // namespace gdc
}  // namespace igcn

using namespace igcn;

/* ppr (B,R,R) f64 -> per subject exactly R*k edges in row-major COO order: edge_src / edge_dst (B*R*k) LOCAL node ids, edge_attr f32.
 * (A column whose kept weight rounds to 0 in fp32 would emit fewer edges; this cannot happen for PPR matrices and the slot is then
 * left as (0, 0, 0) -- igcn_collate_csr treats it as a zero-weight self loop of node 0.) */
extern "C" int igcn_gdc_topk_emit(const double* ppr, int64_t B, int64_t R, int64_t k, int32_t* edge_src, int32_t* edge_dst, float* edge_attr,
                                  void* stream) {
    IGCN_REQUIRE(B >= 0 && R > 0 && k >= 1 && k <= gdc::KMAX && k <= R, IGCN_ERR_BAD_ARG, "gdc_topk_emit: bad sizes (k must be 1..%d)", gdc::KMAX);
    if (B == 0) return IGCN_OK;
    IGCN_REQUIRE(ppr && edge_src && edge_dst && edge_attr, IGCN_ERR_BAD_ARG, "gdc_topk_emit: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    cudaMemsetAsync(edge_src, 0, sizeof(int32_t) * (size_t)(B * R * k), st);
    cudaMemsetAsync(edge_dst, 0, sizeof(int32_t) * (size_t)(B * R * k), st);
    cudaMemsetAsync(edge_attr, 0, sizeof(float) * (size_t)(B * R * k), st);
    const size_t smem = sizeof(int) * (size_t)(3 * R + 1 + 2 * R * k);
    auto kern = gdc::topk_emit_kernel<double>;
    int rc = allow_smem(kern, smem, "gdc_topk_emit");
    if (rc) return rc;
    int64_t grid = (int64_t)sm_count() * 4;
    if (grid > B) grid = B;
    kern<<<(unsigned)grid, 256, smem, st>>>(ppr, (int)B, (int)R, (int)k, edge_src, edge_dst, edge_attr);
    IGCN_CHECK_LAUNCH("gdc_topk_emit");
    return IGCN_OK;
}

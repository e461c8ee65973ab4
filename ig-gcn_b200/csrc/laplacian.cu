// The consistency loss tr(s^T (D - W) s) / B^2 of kernel/sgcn_img_snp.py:183-196 without its cancellation problem.
//
// W = RBF similarity of the subjects' tsne_fdim rows (util/image_cluster.py:15-31: exp(-gamma ||t_i - t_j||^2)), D = diag of its
// row sums, L = D - W.  Round 1 built L with eager torch (cdist, exp, diag, sum: ~10 launches) and computed T = L s as ONE
// tensor-core product; that product cancels: L 1 = 0, W is nearly constant (0.86 +- 0.05 for ADNI-shaped inputs), so T is the
// small difference of two large sums and the 2^-22 operand error of the 3 x TF32 scheme showed up as ~1e-4 in every gradient
// downstream of out_z (measured at B = 256 against the fp64 oracle).  The fix is algebra, not precision:
//     L s = L (s - 1 m^T)   for any row vector m  (L 1 = 0)   ->   centre the columns:  s' = s - mean_rows(s)
//     T   = d .* s' - W s'                                      ->   the dominant diagonal term is an exact fp32 product,
// and the tensor cores only see W s', a sum of terms that already cancel to something small.  Value <s', T> and gradient 2 T / B^2
// are unchanged.  Kernels here: the similarity matrix and its row sums, the column means, and the finishing pass
// T = d .* s' - U (U = W s' from igcn_tc_gemm) fused with the partial sums of <s', T>.
#include "common.cuh"

namespace igcn {
namespace lap {

constexpr int TB = 64;      // similarity tile
// W[i][j] = exp(-gamma * ||t_i - t_j||^2), 64 x 64 tiles, 256 threads (4 x 4 per thread), k-chunks of 32 staged in shared memory
__global__ void __launch_bounds__(256) rbf_kernel(const float* __restrict__ t, int B, int R, float gamma, float* __restrict__ W) {
    IGCN_PDL_SYNC();
    __shared__ float ta[32][TB + 1], tb[32][TB + 1];
    const int i0 = blockIdx.y * TB, j0 = blockIdx.x * TB;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    float acc[4][4];
#pragma unroll
    for (int p = 0; p < 4; ++p)
#pragma unroll
        for (int q = 0; q < 4; ++q) acc[p][q] = 0.f;
    for (int k0 = 0; k0 < R; k0 += 32) {
        for (int idx = threadIdx.x; idx < 32 * TB; idx += 256) {
            const int r = idx >> 5, k = idx & 31;              // coalesced along k
            ta[k][r] = (i0 + r < B && k0 + k < R) ? t[(int64_t)(i0 + r) * R + k0 + k] : 0.f;
            tb[k][r] = (j0 + r < B && k0 + k < R) ? t[(int64_t)(j0 + r) * R + k0 + k] : 0.f;
        }
        __syncthreads();
#pragma unroll 8
        for (int k = 0; k < 32; ++k) {
            float a[4], b[4];
#pragma unroll
            for (int p = 0; p < 4; ++p) {
                a[p] = ta[k][ty * 4 + p];
                b[p] = tb[k][tx * 4 + p];
            }
#pragma unroll
            for (int p = 0; p < 4; ++p)
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const float d = a[p] - b[q];                 // the squared distance directly: no ||a||^2 + ||b||^2 - 2ab cancellation
                    acc[p][q] = fmaf(d, d, acc[p][q]);
                }
        }
        __syncthreads();
    }
#pragma unroll
    for (int p = 0; p < 4; ++p) {
        const int i = i0 + ty * 4 + p;
        if (i >= B) continue;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int j = j0 + tx * 4 + q;
            if (j < B) W[(int64_t)i * B + j] = expf(-gamma * acc[p][q]);
        }
    }
}
// d[i] = sum_j W[i][j] : warp per row, fixed order
__global__ void __launch_bounds__(256) rowsum_kernel(const float* __restrict__ W, int B, float* __restrict__ d) {
    IGCN_PDL_SYNC();
    const int lane = threadIdx.x & 31;
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (row >= B) return;
    float s = 0.f;
    for (int j = lane; j < B; j += 32) s += W[(int64_t)row * B + j];
    s = warp_sum(s);
    if (lane == 0) d[row] = s;
}
// m[g][c] = mean over the B rows of group g of s[(g*B + i)][c] : block = 32 columns x 8 row-lanes, fixed order
__global__ void __launch_bounds__(256) colmean_kernel(const float* __restrict__ s, int B, int D, float* __restrict__ m) {
    IGCN_PDL_SYNC();
    __shared__ float sm[8][33];
    const int g = blockIdx.y, lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int c = blockIdx.x * 32 + lane;
    float acc = 0.f;
    if (c < D) {
        const float* sp = s + (int64_t)g * B * D + c;
#pragma unroll 8
        for (int i = w; i < B; i += 8) acc += sp[(int64_t)i * D];          // unrolled: 8 loads in flight per thread
    }
    sm[w][lane] = acc;
    __syncthreads();
    if (w == 0 && c < D) {
        float tsum = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) tsum += sm[k][lane];
        m[(int64_t)g * D + c] = tsum / (float)B;
    }
}
// T = d[i] * (s - m) - U ; partial sums of <s - m, T>   (U = W (s - m) or NULL when W is all ones: then W s' = 0 and d = B)
__global__ void __launch_bounds__(256) finish_kernel(const float* __restrict__ s, const float* __restrict__ m, const float* __restrict__ d,
                                                     const float* __restrict__ U, int B, int D, int groups, float dconst,
                                                     float* __restrict__ T, float* __restrict__ partials) {
    IGCN_PDL_SYNC();
    __shared__ float sm[8];
    float acc = 0.f;
    for (int row = blockIdx.x; row < groups * B; row += gridDim.x) {        // a CTA walks whole rows: no per-element divisions
        const int g = row / B, i = row - g * B;
        const float di = d ? d[i] : dconst;
        const float* mp = m + (int64_t)g * D;
        const int64_t base = (int64_t)row * D;
        for (int c0 = 0; c0 < D; c0 += 4 * 256) {               // four columns per thread in flight (12 loads) before the first store
            float sv[4], mv[4], uv[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int c = c0 + threadIdx.x + u * 256;
                sv[u] = c < D ? s[base + c] : 0.f;
                mv[u] = c < D ? mp[c] : 0.f;
                uv[u] = (U && c < D) ? U[base + c] : 0.f;
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int c = c0 + threadIdx.x + u * 256;
                if (c < D) {
                    const float sc = sv[u] - mv[u];
                    const float tv = di * sc - uv[u];
                    T[base + c] = tv;
                    acc = fmaf(sc, tv, acc);
                }
            }
        }
    }
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        float tsum = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) tsum += sm[k];
        partials[blockIdx.x] = tsum;
    }
}
__global__ void __launch_bounds__(256) sum_scale_kernel(const float* __restrict__ partials, int n, float scale, float* __restrict__ out) {
    IGCN_PDL_SYNC();
    __shared__ double sm[8];
    double acc = 0.0;
    for (int i = threadIdx.x; i < n; i += 256) acc += (double)partials[i];
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double tsum = 0.0;
        for (int k = 0; k < 8; ++k) tsum += sm[k];
        out[0] = (float)(tsum * (double)scale);
    }
}

}  // namespace lap
}  // namespace igcn

using namespace igcn;

/* W (B,B) = exp(-gamma ||t_i - t_j||^2) for t (B,R) and its row sums d (B). */
extern "C" int igcn_rbf_similarity(const float* t, int64_t B, int64_t R, double gamma, float* W, float* d, void* stream) {
    IGCN_REQUIRE(B >= 0 && R > 0, IGCN_ERR_BAD_ARG, "rbf_similarity: bad size");
    if (B == 0) return IGCN_OK;
    IGCN_REQUIRE(t && W && d, IGCN_ERR_BAD_ARG, "rbf_similarity: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    const int tiles = (int)((B + lap::TB - 1) / lap::TB);
    igcn::launch_k(lap::rbf_kernel, dim3(dim3(tiles, tiles)), dim3(256), 0, st, t, (int)B, (int)R, (float)gamma, W);
    IGCN_CHECK_LAUNCH("rbf_similarity");
    igcn::launch_k(lap::rowsum_kernel, dim3((int)((B + 7) / 8)), dim3(256), 0, st, W, (int)B, d);
    IGCN_CHECK_LAUNCH("rbf_rowsum");
    return IGCN_OK;
}

/* m (groups, D) = column means of every group of B rows of s (groups*B, D). */
extern "C" int igcn_col_mean(const float* s, int64_t B, int64_t D, int64_t groups, float* m, void* stream) {
    IGCN_REQUIRE(B > 0 && D > 0 && groups > 0, IGCN_ERR_BAD_ARG, "col_mean: bad size");
    IGCN_REQUIRE(s && m, IGCN_ERR_BAD_ARG, "col_mean: null pointer");
    igcn::launch_k(lap::colmean_kernel, dim3(dim3((unsigned)((D + 31) / 32), (unsigned)groups)), dim3(256), 0, (cudaStream_t)stream, s, (int)B, (int)D, m);
    IGCN_CHECK_LAUNCH("col_mean");
    return IGCN_OK;
}

/* T = d .* (s - m) - U and out = scale * <s - m, T>;  d == NULL: d_i = d_const;  U == NULL: U = 0.  n_partials = igcn_reduce_blocks(n). */
extern "C" int igcn_laplacian_finish(const float* s, const float* m, const float* d, const float* U, int64_t B, int64_t D, int64_t groups,
                                     double d_const, double scale, float* T, float* partials, int64_t n_partials, float* out, void* stream) {
    IGCN_REQUIRE(B > 0 && D > 0 && groups > 0 && n_partials > 0, IGCN_ERR_BAD_ARG, "laplacian_finish: bad size");
    IGCN_REQUIRE(s && m && T && partials && out, IGCN_ERR_BAD_ARG, "laplacian_finish: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    igcn::launch_k(lap::finish_kernel, dim3((unsigned)n_partials), dim3(256), 0, st, s, m, d, U, (int)B, (int)D, (int)groups, (float)d_const, T, partials);
    IGCN_CHECK_LAUNCH("laplacian_finish");
    igcn::launch_k(lap::sum_scale_kernel, dim3(1), dim3(256), 0, st, partials, (int)n_partials, (float)scale, out);
    IGCN_CHECK_LAUNCH("laplacian_sum");
    return IGCN_OK;
}

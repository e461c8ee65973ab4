// Imaging-SNP fusion heads: relu([X1 | X2 | X3] W^T + b) without materialising the concatenation
// (reference: kernel/sgcn_img_snp.py:286-305 -- out_lin = cat(out_z, latent); lin1; feat4regr = cat(out_lin, img_feat);
// lin1_regr -- two tall-skinny fp32 GEMMs, K = R*L*H+32 (+R*F0) = 2912 / 3182 at R=90, N = 64, M = batch).
//
// ncu on the cuBLAS path at M=256 showed each of these GEMMs running on 4 CTAs (128x32 tiles over a 2912-long K loop,
// 38 us) -- a latency-bound shape, not a tensor-core shape.  Here the K range is split over ~one CTA per SM, the
// partial tiles are reduced in a fixed order (deterministic) together with bias + ReLU, and the three inputs are read in
// place.  fp32 FFMA with 4x4 register tiles; the backward is two more tiled kernels (dW = gZ^T X, dX_s = gZ W_s).
#include "common.cuh"

namespace igcn {

constexpr int TM = 64, TN = 64, TK = 32;

struct CatSrc {
    const float* p[3];
    int w[3];       // widths
    int ld[3];      // row strides
};

__device__ __forceinline__ float cat_load(const CatSrc& s, int row, int k) {
    if (k < s.w[0]) return s.p[0][(int64_t)row * s.ld[0] + k];
    k -= s.w[0];
    if (k < s.w[1]) return s.p[1][(int64_t)row * s.ld[1] + k];
    k -= s.w[1];
    return s.p[2][(int64_t)row * s.ld[2] + k];
}

// C(64x64) += A(64 x TK) * B(TK x 64); As/Bs are k-major: As[k][m], Bs[k][n]; thread (ty,tx) owns a 4x4 block
#define IGCN_TILE_MMA(As, Bs, acc)                                                      \
    _Pragma("unroll") for (int kk = 0; kk < TK; ++kk) {                                  \
        const float4 av = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);             \
        const float4 bv = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);             \
        const float a_[4] = {av.x, av.y, av.z, av.w};                                    \
        const float b_[4] = {bv.x, bv.y, bv.z, bv.w};                                    \
        _Pragma("unroll") for (int i_ = 0; i_ < 4; ++i_)                                 \
            _Pragma("unroll") for (int j_ = 0; j_ < 4; ++j_) acc[i_][j_] = fmaf(a_[i_], b_[j_], acc[i_][j_]); \
    }

// ---- forward, split-K partials: part[s][m][n] = sum_{k in chunk s} X[m][k] W[n][k] -------------------------------------
__global__ void __launch_bounds__(256) cat_linear_fwd_partial(CatSrc src, const float* __restrict__ W, int M, int N, int K, int kchunk,
                                                              float* __restrict__ part) {
    __shared__ __align__(16) float As[TK][TM + 4];
    __shared__ __align__(16) float Bs[TK][TN + 4];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int m0 = blockIdx.x * TM, n0 = blockIdx.y * TN, s = blockIdx.z;
    const int k_begin = s * kchunk, k_end = min(K, k_begin + kchunk);
    float acc[4][4] = {};
    // register double buffering: the next tile's global loads are in flight while the current tile is multiplied
    float ra[8], rb[8];
    auto fetch = [&](int k0) {
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const int e = tid + q * 256;
            const int r = e / TK, kk = e - r * TK;           // consecutive threads -> consecutive k (coalesced)
            const int m = m0 + r, n = n0 + r, k = k0 + kk;
            ra[q] = (m < M && k < k_end) ? cat_load(src, m, k) : 0.f;
            rb[q] = (n < N && k < k_end) ? W[(int64_t)n * K + k] : 0.f;
        }
    };
    if (k_begin < k_end) fetch(k_begin);
    for (int k0 = k_begin; k0 < k_end; k0 += TK) {
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const int e = tid + q * 256;
            const int r = e / TK, kk = e - r * TK;
            As[kk][r] = ra[q];
            Bs[kk][r] = rb[q];
        }
        __syncthreads();
        if (k0 + TK < k_end) fetch(k0 + TK);
        IGCN_TILE_MMA(As, Bs, acc)
        __syncthreads();
    }
    float* out = part + (int64_t)s * M * N;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int m = m0 + ty * 4 + i;
        if (m >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + tx * 4 + j;
            if (n < N) out[(int64_t)m * N + n] = acc[i][j];
        }
    }
}

// out[m][n] = act(sum_s part[s][m][n] + b[n])   (fixed summation order)
__global__ void __launch_bounds__(256) cat_linear_fwd_reduce(const float* __restrict__ part, const float* __restrict__ bias, int M, int N,
                                                             int S, int relu, float* __restrict__ out) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (int64_t)M * N) return;
    float v = 0.f;
    for (int s = 0; s < S; ++s) v += part[(int64_t)s * M * N + idx];
    v += bias[idx % N];
    out[idx] = relu ? fmaxf(v, 0.f) : v;
}

// ---- backward -----------------------------------------------------------------------------------------------------------
// gZ[m][n] = gY[m][n] * (Y[m][n] > 0)  is formed on the fly when tiles are loaded.
// dW[n][k] = sum_m gZ[m][n] X[m][k]      tile: 64 n x 64 k, loop over m
__global__ void __launch_bounds__(256) cat_linear_bwd_w(CatSrc src, const float* __restrict__ gY, const float* __restrict__ Y, int relu,
                                                        int M, int N, int K, float* __restrict__ dW, float* __restrict__ db) {
    __shared__ __align__(16) float As[TK][TM + 4];   // As[mm][n]  = gZ[m0+mm][n]
    __shared__ __align__(16) float Bs[TK][TN + 4];   // Bs[mm][k]  = X[m0+mm][k]
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int n0 = blockIdx.y * TM, k0 = blockIdx.x * TN;
    float acc[4][4] = {};
    float dbacc = 0.f;
    float ra[8], rb[8];
    auto fetch = [&](int m0) {
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const int e = tid + q * 256;
            const int mm = e / TM, c = e - mm * TM;          // consecutive threads -> consecutive n / k (coalesced)
            const int m = m0 + mm;
            const int n = n0 + c, k = k0 + c;
            float gz = 0.f;
            if (m < M && n < N) {
                gz = gY[(int64_t)m * N + n];
                if (relu && !(Y[(int64_t)m * N + n] > 0.f)) gz = 0.f;
            }
            ra[q] = gz;
            rb[q] = (m < M && k < K) ? cat_load(src, m, k) : 0.f;
        }
    };
    fetch(0);
    for (int m0 = 0; m0 < M; m0 += TK) {
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const int e = tid + q * 256;
            const int mm = e / TM, c = e - mm * TM;
            As[mm][c] = ra[q];
            Bs[mm][c] = rb[q];
        }
        __syncthreads();
        if (m0 + TK < M) fetch(m0 + TK);
        IGCN_TILE_MMA(As, Bs, acc)
        if (blockIdx.x == 0 && tid < TM) {                   // d bias: column sums of gZ, one thread per n, fixed order
#pragma unroll 8
            for (int mm = 0; mm < TK; ++mm) dbacc += As[mm][tid];
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int n = n0 + ty * 4 + i;
        if (n >= N) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int k = k0 + tx * 4 + j;
            if (k < K) dW[(int64_t)n * K + k] = acc[i][j];
        }
    }
    if (blockIdx.x == 0 && tid < TM && n0 + tid < N) db[n0 + tid] = dbacc;
}

// dX[m][k] = sum_n gZ[m][n] W[n][k]      tile: 64 m x 64 k, loop over n ; scattered to the three sources' gradients
struct CatDst {
    float* p[3];
    int w[3];
    int ld[3];
};
__global__ void __launch_bounds__(256) cat_linear_bwd_x(CatDst dst, const float* __restrict__ gY, const float* __restrict__ Y, int relu,
                                                        const float* __restrict__ W, int M, int N, int K) {
    __shared__ __align__(16) float As[TK][TM + 4];   // As[nn][m] = gZ[m][n0+nn]
    __shared__ __align__(16) float Bs[TK][TN + 4];   // Bs[nn][k] = W[n0+nn][k]
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int m0 = blockIdx.y * TM, k0 = blockIdx.x * TN;
    float acc[4][4] = {};
    float ra[8], rb[8];
    auto fetch = [&](int nb) {
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const int e = tid + q * 256;
            {
                const int r = e / TK, nn = e - r * TK;       // gZ row-major (m, n): consecutive threads -> consecutive n
                const int m = m0 + r, n = nb + nn;
                float gz = 0.f;
                if (m < M && n < N) {
                    gz = gY[(int64_t)m * N + n];
                    if (relu && !(Y[(int64_t)m * N + n] > 0.f)) gz = 0.f;
                }
                ra[q] = gz;
            }
            {
                const int nn = e / TN, c = e - nn * TN;      // W row-major (n, k): consecutive threads -> consecutive k
                const int n = nb + nn, k = k0 + c;
                rb[q] = (n < N && k < K) ? W[(int64_t)n * K + k] : 0.f;
            }
        }
    };
    fetch(0);
    for (int nb = 0; nb < N; nb += TK) {
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const int e = tid + q * 256;
            As[e - (e / TK) * TK][e / TK] = ra[q];
            Bs[e / TN][e - (e / TN) * TN] = rb[q];
        }
        __syncthreads();
        if (nb + TK < N) fetch(nb + TK);
        IGCN_TILE_MMA(As, Bs, acc)
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int m = m0 + ty * 4 + i;
        if (m >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int k = k0 + tx * 4 + j;
            if (k >= K) continue;
            int sidx = 0;
            if (k >= dst.w[0]) {
                k -= dst.w[0];
                sidx = 1;
                if (k >= dst.w[1]) {
                    k -= dst.w[1];
                    sidx = 2;
                }
            }
            float* p = dst.p[sidx];
            if (p) p[(int64_t)m * dst.ld[sidx] + k] = acc[i][j];
        }
    }
}

static int cat_check(const char* who, const float* const* x, const int64_t* w, const int64_t* ld, int64_t M, int64_t N, int64_t K,
                     bool need_all) {
    IGCN_REQUIRE(M >= 0 && N > 0 && K > 0, IGCN_ERR_BAD_ARG, "%s: bad size", who);
    IGCN_REQUIRE(w[0] >= 0 && w[1] >= 0 && w[2] >= 0 && w[0] + w[1] + w[2] == K, IGCN_ERR_BAD_ARG, "%s: source widths must add up to K", who);
    for (int i = 0; i < 3; ++i) {
        IGCN_REQUIRE(w[i] == 0 || ld[i] >= w[i], IGCN_ERR_BAD_ARG, "%s: row stride of source %d smaller than its width", who, i);
        IGCN_REQUIRE(!need_all || w[i] == 0 || x[i], IGCN_ERR_BAD_ARG, "%s: null source %d", who, i);
    }
    return IGCN_OK;
}

}  // namespace igcn

using namespace igcn;

extern "C" int64_t igcn_cat_linear_splits(int64_t M, int64_t N, int64_t K) {
    const int64_t tiles = ((M + TM - 1) / TM) * ((N + TN - 1) / TN);
    int64_t S = (sm_count() + tiles - 1) / tiles;            // about one CTA per SM
    const int64_t max_s = (K + TK - 1) / TK;
    if (S > max_s) S = max_s;
    if (S < 1) S = 1;
    return S;
}

extern "C" int igcn_cat_linear_fwd(const float* x0, const float* x1, const float* x2, const int64_t* host_widths, const int64_t* host_strides,
                                   const float* W, const float* bias, int64_t M, int64_t N, int64_t K, int64_t relu,
                                   float* partials, int64_t S, float* out, void* stream) {
    const float* xs[3] = {x0, x1, x2};
    const int64_t *widths = host_widths, *strides = host_strides;
    IGCN_REQUIRE(widths && strides, IGCN_ERR_BAD_ARG, "cat_linear_fwd: null host arrays");
    int rc = cat_check("cat_linear_fwd", xs, widths, strides, M, N, K, true);
    if (rc) return rc;
    IGCN_REQUIRE(W && bias && partials && out, IGCN_ERR_BAD_ARG, "cat_linear_fwd: null pointer");
    IGCN_REQUIRE(S == igcn_cat_linear_splits(M, N, K), IGCN_ERR_BAD_ARG, "cat_linear_fwd: S must be igcn_cat_linear_splits()");
    if (M == 0) return IGCN_OK;
    CatSrc src;
    for (int i = 0; i < 3; ++i) { src.p[i] = xs[i]; src.w[i] = (int)widths[i]; src.ld[i] = (int)strides[i]; }
    const int kchunk = (int)(((K + S - 1) / S + TK - 1) / TK * TK);
    cudaStream_t st = (cudaStream_t)stream;
    dim3 grid((unsigned)((M + TM - 1) / TM), (unsigned)((N + TN - 1) / TN), (unsigned)S);
    cat_linear_fwd_partial<<<grid, 256, 0, st>>>(src, W, (int)M, (int)N, (int)K, kchunk, partials);
    IGCN_CHECK_LAUNCH("cat_linear_fwd_partial");
    cat_linear_fwd_reduce<<<(unsigned)((M * N + 255) / 256), 256, 0, st>>>(partials, bias, (int)M, (int)N, (int)S, (int)relu, out);
    IGCN_CHECK_LAUNCH("cat_linear_fwd_reduce");
    return IGCN_OK;
}

extern "C" int igcn_cat_linear_bwd(const float* x0, const float* x1, const float* x2, const int64_t* host_widths, const int64_t* host_strides,
                                   const float* W, const float* out, const float* g_out, int64_t M, int64_t N, int64_t K, int64_t relu,
                                   float* dx0, float* dx1, float* dx2, const int64_t* host_dstrides, float* dW, float* db, void* stream) {
    const float* xs[3] = {x0, x1, x2};
    const int64_t *widths = host_widths, *strides = host_strides, *dstrides = host_dstrides;
    IGCN_REQUIRE(widths && strides, IGCN_ERR_BAD_ARG, "cat_linear_bwd: null host arrays");
    int rc = cat_check("cat_linear_bwd", xs, widths, strides, M, N, K, true);
    if (rc) return rc;
    IGCN_REQUIRE(W && out && g_out && dstrides && ((dW == nullptr) == (db == nullptr)), IGCN_ERR_BAD_ARG, "cat_linear_bwd: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    if (M == 0) {
        if (dW) {
            cudaMemsetAsync(dW, 0, sizeof(float) * N * K, st);
            cudaMemsetAsync(db, 0, sizeof(float) * N, st);
        }
        return IGCN_OK;
    }
    CatSrc src;
    CatDst dst;
    float* dxs[3] = {dx0, dx1, dx2};
    for (int i = 0; i < 3; ++i) {
        src.p[i] = xs[i]; src.w[i] = (int)widths[i]; src.ld[i] = (int)strides[i];
        dst.p[i] = dxs[i]; dst.w[i] = (int)widths[i]; dst.ld[i] = (int)dstrides[i];
    }
    if (dW) {   // dW == db == NULL: weight gradient not wanted
        dim3 gw((unsigned)((K + TN - 1) / TN), (unsigned)((N + TM - 1) / TM));
        cat_linear_bwd_w<<<gw, 256, 0, st>>>(src, g_out, out, (int)relu, (int)M, (int)N, (int)K, dW, db);
        IGCN_CHECK_LAUNCH("cat_linear_bwd_w");
    }
    if (dx0 || dx1 || dx2) {
        dim3 gx((unsigned)((K + TN - 1) / TN), (unsigned)((M + TM - 1) / TM));
        cat_linear_bwd_x<<<gx, 256, 0, st>>>(dst, g_out, out, (int)relu, W, (int)M, (int)N, (int)K);
        IGCN_CHECK_LAUNCH("cat_linear_bwd_x");
    }
    return IGCN_OK;
}

// Fusion heads at the reference's batch sizes: act([x0 | x1 | x2] W^T + b) with N <= 64 outputs and a few hundred rows
// (kernel/sgcn_img_snp.py:286-305, lin1 / lin1_regr: K = R*L*H + 32 (+ R*F0), N = 64, M = 2 x batch), forward and both backward
// products, on warp-level tensor cores (mma.sync m16n8k8 TF32, split in three, fp32 accumulate -- the scheme of mma_util.cuh).
//
// Why not the tcgen05 kernel (tc_gemm.cu) here: at M = 512 that path is split pass -> 128-row UMMA tiles with split-K -> reduce
// pass; CUPTI on the benchmarked step (profiles/r2_bench_config2_l.json): 11.6 + 8.1 + 6.2 us forward and 11.6 + 27.7 us per
// backward product for 0.2 GFLOP each -- 26 % of the step's critical path, all of it launch latency, operand copies and tile
// quantisation (4 row tiles for 148 SMs).  These kernels read the fp32 operands once, in place (the concatenation, the row
// repetition of a shared source and the ReLU mask are index arithmetic), split hi / lo in registers, and size their grids by
// 16-row tiles, so one launch per product is enough.  tc_gemm keeps the large shapes (config 4: M = 8 192).
//
// Fragment layouts (PTX ISA, m16n8k8 .tf32; g = lane / 4, t = lane % 4):
//   A (16 x 8): a0 (g, t) a1 (g+8, t) a2 (g, t+4) a3 (g+8, t+4);  B (8 x 8): b0 (k = t, n = g) b1 (k = t+4, n = g);
//   C (16 x 8): c0 (g, 2t) c1 (g, 2t+1) c2 (g+8, 2t) c3 (g+8, 2t+1).
#include <stdlib.h>

#include "common.cuh"
#include "mma_util.cuh"

namespace igcn {
namespace catlin {

using namespace igcn::mmau;

struct Src {
    const float* p[3];   // (rows_i, w_i), row stride ld_i; row m of the product reads row m % rows_i
    int w[3], ld[3], rows[3];
};
struct Dst {
    float* p[3];         // (M, w_i) input gradients, row stride ld_i; NULL = not needed
    int ld[3];
};

// the concatenated list of 8-wide column steps: segment s contributes ceil(w_s / 8) steps (tails are zero padded, steps never
// straddle two sources); seg = 3 is the bias column of the weight-gradient product (a column of ones)
struct Steps {
    int n[4], total;
};
__device__ __forceinline__ Steps make_steps(const Src& s, bool ones) {
    Steps st;
    st.n[0] = (s.w[0] + 7) >> 3; st.n[1] = (s.w[1] + 7) >> 3; st.n[2] = (s.w[2] + 7) >> 3; st.n[3] = ones ? 1 : 0;
    st.total = st.n[0] + st.n[1] + st.n[2] + st.n[3];
    return st;
}
__device__ __forceinline__ void locate(const Steps& st, const Src& s, int i, int& seg, int& k0, int& off) {
    seg = 0; off = 0;
    if (i >= st.n[0]) { i -= st.n[0]; off += s.w[0]; seg = 1;
        if (i >= st.n[1]) { i -= st.n[1]; off += s.w[1]; seg = 2;
            if (i >= st.n[2]) { i -= st.n[2]; off += s.w[2]; seg = 3; } } }
    k0 = 8 * i;
}
// field of segment `seg` without dynamic indexing of the kernel-parameter struct (which would be copied to local memory)
#define CATLIN_SEL(arr, seg) ((seg) == 0 ? (arr)[0] : ((seg) == 1 ? (arr)[1] : (arr)[2]))

// ---- forward: one CTA per 16 output rows, the warps split the K steps, fixed-order reduction through shared memory ------------------
constexpr int FW_WARPS = 16, FW_LD = 72;     // 72: conflict-free float2 stores of the accumulator fragments

__global__ void __launch_bounds__(32 * FW_WARPS) catlin_fwd_kernel(Src s, const float* __restrict__ W, const float* __restrict__ bias, int M,
                                                                   int N, int K, int relu, float* __restrict__ Y) {
    extern __shared__ __align__(16) float red[];       // [FW_WARPS][16][FW_LD]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;
    const int m0 = blockIdx.x * 16, r0 = m0 + g, r1 = r0 + 8;
    const Steps st = make_steps(s, false);
    const int per = (st.total + FW_WARPS - 1) / FW_WARPS, lo = warp * per, hi = min(st.total, lo + per);
    float acc[8][4];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f;
#pragma unroll 2
    for (int i = lo; i < hi; ++i) {
        int seg, k0, off;
        locate(st, s, i, seg, k0, off);
        const float* xp = CATLIN_SEL(s.p, seg);
        const int w = CATLIN_SEL(s.w, seg), ld = CATLIN_SEL(s.ld, seg), rows = CATLIN_SEL(s.rows, seg);
        const int ka = k0 + t, kb = ka + 4;
        const bool va = ka < w, vb = kb < w;
        const float* x0 = xp + (int64_t)(r0 % rows) * ld;
        const float* x1 = xp + (int64_t)(r1 % rows) * ld;
        uint32_t ah[4], al[4];
        split((r0 < M && va) ? x0[ka] : 0.f, ah[0], al[0]);
        split((r1 < M && va) ? x1[ka] : 0.f, ah[1], al[1]);
        split((r0 < M && vb) ? x0[kb] : 0.f, ah[2], al[2]);
        split((r1 < M && vb) ? x1[kb] : 0.f, ah[3], al[3]);
        const float* wp = W + off;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int n = 8 * j + g;
            if (8 * j < N) {
                uint32_t bh0, bl0, bh1, bl1;
                split((n < N && va) ? wp[(int64_t)n * K + ka] : 0.f, bh0, bl0);
                split((n < N && vb) ? wp[(int64_t)n * K + kb] : 0.f, bh1, bl1);
                mma_k8(acc[j], al[0], al[1], al[2], al[3], bh0, bh1);
                mma_k8(acc[j], ah[0], ah[1], ah[2], ah[3], bl0, bl1);
                mma_k8(acc[j], ah[0], ah[1], ah[2], ah[3], bh0, bh1);
            }
        }
    }
    float* rw = red + warp * 16 * FW_LD;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        *reinterpret_cast<float2*>(rw + g * FW_LD + 8 * j + 2 * t) = make_float2(acc[j][0], acc[j][1]);
        *reinterpret_cast<float2*>(rw + (g + 8) * FW_LD + 8 * j + 2 * t) = make_float2(acc[j][2], acc[j][3]);
    }
    __syncthreads();
    for (int idx = tid; idx < 16 * 64; idx += 32 * FW_WARPS) {
        const int r = idx >> 6, c = idx & 63;
        if (m0 + r < M && c < N) {
            float v = 0.f;
#pragma unroll
            for (int wv = 0; wv < FW_WARPS; ++wv) v += red[(wv * 16 + r) * FW_LD + c];
            if (bias) v += bias[c];
            if (relu) v = fmaxf(v, 0.f);
            Y[(int64_t)(m0 + r) * N + c] = v;
        }
    }
}

// gY through the ReLU mask of the forward output
__device__ __forceinline__ float masked(const float* __restrict__ gY, const float* __restrict__ out, int64_t i) {
    const float v = gY[i];
    return (out && !(out[i] > 0.f)) ? 0.f : v;
}

// ---- input gradient: dX[m][k] = sum_n dZ[m][n] W[n][k]; one CTA per (16-row tile, slice of the 8-column output tiles) ---------------
constexpr int DX_WARPS = 8;

__global__ void __launch_bounds__(32 * DX_WARPS) catlin_dx_kernel(Src s, Dst d, const float* __restrict__ W, const float* __restrict__ gY,
                                                                  const float* __restrict__ out, int M, int N, int K) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;
    const int m0 = blockIdx.x * 16, r0 = m0 + g, r1 = r0 + 8;
    uint32_t ah[8][4], al[8][4];                        // dZ rows r0, r1 over the N <= 64 contraction steps
#pragma unroll
    for (int ks = 0; ks < 8; ++ks) {
        const int na = 8 * ks + t, nb = na + 4;
        split((r0 < M && na < N) ? masked(gY, out, (int64_t)r0 * N + na) : 0.f, ah[ks][0], al[ks][0]);
        split((r1 < M && na < N) ? masked(gY, out, (int64_t)r1 * N + na) : 0.f, ah[ks][1], al[ks][1]);
        split((r0 < M && nb < N) ? masked(gY, out, (int64_t)r0 * N + nb) : 0.f, ah[ks][2], al[ks][2]);
        split((r1 < M && nb < N) ? masked(gY, out, (int64_t)r1 * N + nb) : 0.f, ah[ks][3], al[ks][3]);
    }
    const Steps st = make_steps(s, false);
    for (int i = blockIdx.y * DX_WARPS + warp; i < st.total; i += gridDim.y * DX_WARPS) {
        int seg, c0, off;
        locate(st, s, i, seg, c0, off);
        float* dp = CATLIN_SEL(d.p, seg);
        if (!dp) continue;
        const int w = CATLIN_SEL(s.w, seg);
        const bool vc = c0 + g < w;                     // B column n = g -> output column c0 + g
        const float* wp = W + off + c0 + g;
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {
            if (8 * ks < N) {
                const int na = 8 * ks + t, nb = na + 4;
                uint32_t bh0, bl0, bh1, bl1;
                split((vc && na < N) ? wp[(int64_t)na * K] : 0.f, bh0, bl0);
                split((vc && nb < N) ? wp[(int64_t)nb * K] : 0.f, bh1, bl1);
                mma_k8(acc, al[ks][0], al[ks][1], al[ks][2], al[ks][3], bh0, bh1);
                mma_k8(acc, ah[ks][0], ah[ks][1], ah[ks][2], ah[ks][3], bl0, bl1);
                mma_k8(acc, ah[ks][0], ah[ks][1], ah[ks][2], ah[ks][3], bh0, bh1);
            }
        }
        const int c = c0 + 2 * t, ld = CATLIN_SEL(d.ld, seg);
        if (r0 < M) {
            if (c < w) dp[(int64_t)r0 * ld + c] = acc[0];
            if (c + 1 < w) dp[(int64_t)r0 * ld + c + 1] = acc[1];
        }
        if (r1 < M) {
            if (c < w) dp[(int64_t)r1 * ld + c] = acc[2];
            if (c + 1 < w) dp[(int64_t)r1 * ld + c + 1] = acc[3];
        }
    }
}

// ---- weight gradient: dW[n][k] = sum_m dZ[m][n] X[m][k], db[n] = sum_m dZ[m][n]; one CTA per GC 8-column tiles of [X | 1], the warps
//      split the rows (the contraction) and are summed in warp order ------------------------------------------------------------------
constexpr int DW_WARPS = 8, GC = 4, DW_LD = 8 * GC + 8;      // 40: conflict-free float2 stores

__global__ void __launch_bounds__(32 * DW_WARPS) catlin_dw_kernel(Src s, const float* __restrict__ gY, const float* __restrict__ out, int M,
                                                                  int N, int K, float* __restrict__ dW, float* __restrict__ db) {
    extern __shared__ __align__(16) float red[];       // [DW_WARPS][64][DW_LD]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;
    const Steps st = make_steps(s, true);
    const int ksteps = (M + 7) >> 3, per = (ksteps + DW_WARPS - 1) / DW_WARPS;
    const int ks_lo = warp * per, ks_hi = min(ksteps, ks_lo + per);
    int seg[GC], c0[GC], off[GC], ldq[GC], rowsq[GC];
    const float* xq[GC];                                 // column c0 + g of the tile's source, NULL = nothing to read
#pragma unroll
    for (int q = 0; q < GC; ++q) {
        const int i = blockIdx.x * GC + q;
        if (i < st.total) locate(st, s, i, seg[q], c0[q], off[q]);
        else { seg[q] = -1; c0[q] = 0; off[q] = 0; }
        xq[q] = nullptr; ldq[q] = 0; rowsq[q] = 1;
        if (seg[q] >= 0 && seg[q] < 3 && c0[q] + g < CATLIN_SEL(s.w, seg[q])) {
            xq[q] = CATLIN_SEL(s.p, seg[q]) + c0[q] + g;
            ldq[q] = CATLIN_SEL(s.ld, seg[q]);
            rowsq[q] = CATLIN_SEL(s.rows, seg[q]);
        }
    }
    float acc[4][GC][4];
#pragma unroll
    for (int mt = 0; mt < 4; ++mt)
#pragma unroll
        for (int q = 0; q < GC; ++q) acc[mt][q][0] = acc[mt][q][1] = acc[mt][q][2] = acc[mt][q][3] = 0.f;
    for (int ks = ks_lo; ks < ks_hi; ++ks) {
        const int ma = 8 * ks + t, mb = ma + 4;          // contraction rows of this step
        uint32_t bh[GC][2], bl[GC][2];
#pragma unroll
        for (int q = 0; q < GC; ++q) {
            float b0 = 0.f, b1 = 0.f;
            if (seg[q] == 3) {                          // the bias column: ones in column 0 of the tile
                b0 = (g == 0 && ma < M) ? 1.f : 0.f;
                b1 = (g == 0 && mb < M) ? 1.f : 0.f;
            } else if (xq[q]) {
                if (ma < M) b0 = xq[q][(int64_t)(ma % rowsq[q]) * ldq[q]];
                if (mb < M) b1 = xq[q][(int64_t)(mb % rowsq[q]) * ldq[q]];
            }
            split(b0, bh[q][0], bl[q][0]);
            split(b1, bh[q][1], bl[q][1]);
        }
#pragma unroll
        for (int mt = 0; mt < 4; ++mt) {
            if (16 * mt < N) {
                const int na = 16 * mt + g, nb = na + 8;   // A = dZ^T: rows = output feature n, k = batch row
                uint32_t ah[4], al[4];
                split((ma < M && na < N) ? masked(gY, out, (int64_t)ma * N + na) : 0.f, ah[0], al[0]);
                split((ma < M && nb < N) ? masked(gY, out, (int64_t)ma * N + nb) : 0.f, ah[1], al[1]);
                split((mb < M && na < N) ? masked(gY, out, (int64_t)mb * N + na) : 0.f, ah[2], al[2]);
                split((mb < M && nb < N) ? masked(gY, out, (int64_t)mb * N + nb) : 0.f, ah[3], al[3]);
#pragma unroll
                for (int q = 0; q < GC; ++q) {
                    mma_k8(acc[mt][q], al[0], al[1], al[2], al[3], bh[q][0], bh[q][1]);
                    mma_k8(acc[mt][q], ah[0], ah[1], ah[2], ah[3], bl[q][0], bl[q][1]);
                    mma_k8(acc[mt][q], ah[0], ah[1], ah[2], ah[3], bh[q][0], bh[q][1]);
                }
            }
        }
    }
    float* rw = red + warp * 64 * DW_LD;
#pragma unroll
    for (int mt = 0; mt < 4; ++mt)
#pragma unroll
        for (int q = 0; q < GC; ++q) {
            *reinterpret_cast<float2*>(rw + (16 * mt + g) * DW_LD + 8 * q + 2 * t) = make_float2(acc[mt][q][0], acc[mt][q][1]);
            *reinterpret_cast<float2*>(rw + (16 * mt + g + 8) * DW_LD + 8 * q + 2 * t) = make_float2(acc[mt][q][2], acc[mt][q][3]);
        }
    __syncthreads();
    for (int idx = tid; idx < 64 * 8 * GC; idx += 32 * DW_WARPS) {
        const int n = idx / (8 * GC), cc = idx - n * (8 * GC), q = cc >> 3, c = cc & 7;
        const int i = blockIdx.x * GC + q;
        if (n >= N || i >= st.total) continue;
        int sg, k0, of;
        locate(st, s, i, sg, k0, of);
        float v = 0.f;
#pragma unroll
        for (int wv = 0; wv < DW_WARPS; ++wv) v += red[(wv * 64 + n) * DW_LD + cc];
        if (sg == 3) {
            if (c == 0 && db) db[n] = v;
        } else if (k0 + c < CATLIN_SEL(s.w, sg)) {
            dW[(int64_t)n * K + of + k0 + c] = v;
        }
    }
}

// Measured on B200 at the benchmarked size (M = 512, K = 2 912 / 3 182, N = 64; CUPTI inside the step, profiles/r2_bench_config2_m.json):
// forward 79 us, dX 47 us, dW 50 us -- SLOWER than the tcgen05 path they were meant to replace (26 / 25 us with its split and
// reduce passes).  The fragments here are loaded straight from global memory: every load instruction touches 8 rows x 16 bytes
// (half a sector each, every sector fetched twice), only ceil(M / 16) = 32 CTAs carry the forward, and each of them streams all
// of W through one L1.  Kept opt-in (IGCN_CATLIN_MMA=1) as a measured negative result; a version that stages 64-row operand tiles
// in shared memory with split-K over the CTAs is the way to do this shape.
static bool enabled() {
    static const bool on = [] { const char* e = getenv("IGCN_CATLIN_MMA"); return e && e[0] == '1'; }();
    return on;
}

static int fill_src(Src& s, const char* who, const float* x0, const float* x1, const float* x2, const int64_t* widths, const int64_t* strides,
                    const int64_t* rows, int64_t M, int64_t K) {
    const float* xs[3] = {x0, x1, x2};
    int64_t wsum = 0;
    for (int i = 0; i < 3; ++i) {
        IGCN_REQUIRE(widths[i] >= 0 && (widths[i] == 0 || (xs[i] && strides[i] >= widths[i] && rows[i] >= 1 && rows[i] <= M)), IGCN_ERR_BAD_ARG,
                     "%s: bad source %d", who, i);
        s.p[i] = xs[i];
        s.w[i] = (int)widths[i];
        s.ld[i] = (int)strides[i];
        s.rows[i] = widths[i] ? (int)rows[i] : 1;
        wsum += widths[i];
    }
    IGCN_REQUIRE(wsum == K, IGCN_ERR_BAD_ARG, "%s: source widths add up to %lld, K = %lld", who, (long long)wsum, (long long)K);
    return IGCN_OK;
}

}  // namespace catlin
}  // namespace igcn

using namespace igcn;

extern "C" int64_t igcn_catlin_mma_supported(int64_t M, int64_t N, int64_t K) {
    return (catlin::enabled() && M >= 1 && M <= 2048 && N >= 16 && N <= 64 && (N % 16) == 0 && K >= 1 && K < (1 << 24)) ? 1 : 0;
}

extern "C" int igcn_catlin_mma_fwd(const float* x0, const float* x1, const float* x2, const int64_t* host_widths, const int64_t* host_strides,
                                   const int64_t* host_rows, const float* W, const float* bias, int64_t M, int64_t N, int64_t K, int64_t relu,
                                   float* out, void* stream) {
    IGCN_REQUIRE(host_widths && host_strides && host_rows && W && out, IGCN_ERR_BAD_ARG, "catlin_mma_fwd: null pointer");
    IGCN_REQUIRE(igcn_catlin_mma_supported(M, N, K), IGCN_ERR_UNSUPPORTED, "catlin_mma_fwd: shape (M=%lld,N=%lld,K=%lld) not supported",
                 (long long)M, (long long)N, (long long)K);
    catlin::Src s;
    int rc = catlin::fill_src(s, "catlin_mma_fwd", x0, x1, x2, host_widths, host_strides, host_rows, M, K);
    if (rc) return rc;
    const size_t smem = sizeof(float) * catlin::FW_WARPS * 16 * catlin::FW_LD;
    if ((rc = allow_smem(catlin::catlin_fwd_kernel, smem, "catlin_mma_fwd"))) return rc;
    catlin::catlin_fwd_kernel<<<(unsigned)((M + 15) / 16), 32 * catlin::FW_WARPS, smem, (cudaStream_t)stream>>>(s, W, bias, (int)M, (int)N, (int)K,
                                                                                                           relu ? 1 : 0, out);
    IGCN_CHECK_LAUNCH("catlin_mma_fwd");
    return IGCN_OK;
}

extern "C" int igcn_catlin_mma_bwd_dx(const int64_t* host_widths, const float* W, const float* out, const float* g_out, int64_t M, int64_t N,
                                      int64_t K, int64_t relu, float* dx0, float* dx1, float* dx2, const int64_t* host_dx_strides, void* stream) {
    IGCN_REQUIRE(host_widths && host_dx_strides && W && g_out && (!relu || out), IGCN_ERR_BAD_ARG, "catlin_mma_bwd_dx: null pointer");
    IGCN_REQUIRE(igcn_catlin_mma_supported(M, N, K), IGCN_ERR_UNSUPPORTED, "catlin_mma_bwd_dx: shape not supported");
    catlin::Src s{};
    catlin::Dst d{};
    float* ds[3] = {dx0, dx1, dx2};
    int64_t wsum = 0;
    for (int i = 0; i < 3; ++i) {
        IGCN_REQUIRE(host_widths[i] >= 0 && (!ds[i] || host_dx_strides[i] >= host_widths[i]), IGCN_ERR_BAD_ARG, "catlin_mma_bwd_dx: bad segment %d", i);
        s.w[i] = (int)host_widths[i]; s.rows[i] = 1;
        d.p[i] = host_widths[i] ? ds[i] : nullptr;
        d.ld[i] = (int)host_dx_strides[i];
        wsum += host_widths[i];
    }
    IGCN_REQUIRE(wsum == K, IGCN_ERR_BAD_ARG, "catlin_mma_bwd_dx: widths do not add up to K");
    if (!dx0 && !dx1 && !dx2) return IGCN_OK;
    const int64_t tiles = (K + 7) / 8 + 3, mt = (M + 15) / 16;
    int64_t gy = (tiles + catlin::DX_WARPS - 1) / catlin::DX_WARPS;          // column slices: enough CTAs for ~4 per SM
    const int64_t want = ((int64_t)sm_count() * 4 + mt - 1) / mt;
    if (gy > want) gy = want;
    if (gy < 1) gy = 1;
    catlin::catlin_dx_kernel<<<dim3((unsigned)mt, (unsigned)gy), 32 * catlin::DX_WARPS, 0, (cudaStream_t)stream>>>(
        s, d, W, g_out, relu ? out : nullptr, (int)M, (int)N, (int)K);
    IGCN_CHECK_LAUNCH("catlin_mma_bwd_dx");
    return IGCN_OK;
}

extern "C" int igcn_catlin_mma_bwd_dw(const float* x0, const float* x1, const float* x2, const int64_t* host_widths, const int64_t* host_strides,
                                      const int64_t* host_rows, const float* out, const float* g_out, int64_t M, int64_t N, int64_t K,
                                      int64_t relu, float* dW, float* db, void* stream) {
    IGCN_REQUIRE(host_widths && host_strides && host_rows && g_out && dW && (!relu || out), IGCN_ERR_BAD_ARG, "catlin_mma_bwd_dw: null pointer");
    IGCN_REQUIRE(igcn_catlin_mma_supported(M, N, K), IGCN_ERR_UNSUPPORTED, "catlin_mma_bwd_dw: shape not supported");
    catlin::Src s;
    int rc = catlin::fill_src(s, "catlin_mma_bwd_dw", x0, x1, x2, host_widths, host_strides, host_rows, M, K);
    if (rc) return rc;
    int64_t steps = 1;
    for (int i = 0; i < 3; ++i) steps += (host_widths[i] + 7) / 8;
    const size_t smem = sizeof(float) * catlin::DW_WARPS * 64 * catlin::DW_LD;
    if ((rc = allow_smem(catlin::catlin_dw_kernel, smem, "catlin_mma_bwd_dw"))) return rc;
    catlin::catlin_dw_kernel<<<(unsigned)((steps + catlin::GC - 1) / catlin::GC), 32 * catlin::DW_WARPS, smem, (cudaStream_t)stream>>>(
        s, g_out, relu ? out : nullptr, (int)M, (int)N, (int)K, dW, db);
    IGCN_CHECK_LAUNCH("catlin_mma_bwd_dw");
    return IGCN_OK;
}

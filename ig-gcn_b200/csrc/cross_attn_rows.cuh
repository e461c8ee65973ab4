// Row-parallel cross attention for the reference's shape (E = 32 embed, M <= 32 GO tokens per subject).
//
// The first kernel (cross_attn.cu) follows nn.MultiheadAttention stage by stage: Q = X Wq^T, S = Q K^T, P = softmax(S),
// O = P V, Y = O Wo^T -- four (R x 32) intermediates per graph in shared memory, two of the five products LDS bound, ~10 block
// barriers per graph (config 4: 2.3 ms forward, 6.7 ms backward, 7-9 % of the FFMA peak).
// With only M = 19 key/value tokens per subject the two projections on the QUERY side fold into the key/value side:
//     S_h = X (scale * K_h Wq_h)^T + scale * K_h bq_h      =: X K'_h^T + c_h        K'_h : (M x 32)
//     Y   = sum_h P_h (V_h Wo_h^T) + bo                    =: sum_h P_h V'_h + bo    V'_h : (M x 32)
// K'_h, V'_h, c_h cost O(M * 16 * 32) per graph and head, and afterwards every QUERY ROW is independent:
//     one thread = one row: x (32 registers) -> 2 x 19 scores -> softmax in registers -> y (32 registers).
// 30 % fewer FMAs than the staged form, no Q / O / dO / dQ intermediates, broadcast LDS.128 operands (1 LDS per 4 FMA), and
// two barriers per graph instead of ten.  The backward is the same row pass (recompute P, dP = dy V'^T, dS, dx = dS K') plus
// two small reductions over the rows of a graph, dV'_h = P_h^T dY and dK'_h = dS_h^T X (4 x 4 register tiles), after which the
// chain back to K, V, the tokens and the six parameter tensors is M x 32 sized.  Deterministic: fixed summation orders, the
// per-CTA parameter-gradient partials are reduced by reduce_partials_kernel.
#pragma once

namespace igcn {
namespace rows {

constexpr int kE = 32;
constexpr int XS = 36;   // staged row stride in floats (144 B: thread-per-row LDS.128 is bank-conflict free); column 32 holds 1.0

__device__ __forceinline__ int round_up(int a, int b) { return (a + b - 1) / b * b; }

struct Geo {
    int gpc;        // graphs per CTA pass
    int nthreads;
    int per_sz;     // floats of per-graph tables
    int mcp;        // P / dS stage row stride (>= M + 1, multiple of 4)
    int T;          // backward: thread groups that split the rows of a graph in the tile reductions (sized for a full pass)
    size_t smem;
};

// forward: per-graph tables = A, K, V (3 M E) | K' (H M E) | V' (H M E) | c (pad4(H M))
static Geo fwd_geo(int R, int M, int H) {
    Geo g;
    g.mcp = 0;
    g.T = 1;
    g.per_sz = 3 * M * kE + 2 * H * M * kE + ((H * M + 3) & ~3);
    int gpc = 224 / R;
    if (gpc < 1) gpc = 1;
    if (gpc > 8) gpc = 8;
    auto smem_of = [&](int n) { return (size_t)4 * (4 * kE * kE + 4 * kE + (size_t)n * g.per_sz + (size_t)n * R * XS) + 16; };
    while (gpc > 1 && smem_of(gpc) > 100 * 1024) --gpc;
    g.gpc = gpc;
    int rows = gpc * R;
    g.nthreads = rows >= 288 ? 288 : ((rows + 31) / 32) * 32;
    g.smem = smem_of(gpc);
    return g;
}

// backward: per-graph tables = A, K, V | K', V' | dK', dV' (each H * mcp * XS: padded rows j and the ones columns) | c | dc
static int bwd_per_sz(int M, int H, int mcp) { return 3 * M * kE + 2 * H * M * kE + 2 * H * mcp * XS + 2 * ((H * M + 3) & ~3) + 2 * M * kE; }
static Geo bwd_geo(int R, int M, int H) {
    Geo g;
    g.mcp = ((M + 1) + 3) & ~3;
    g.per_sz = bwd_per_sz(M, H, g.mcp);
    auto threads_of = [&](int n) {
        const int rows = n * R;
        int nt = rows >= 288 ? 288 : ((rows + 31) / 32) * 32;
        return nt < 128 ? 128 : nt;
    };
    auto split_of = [&](int n) {
        const int tiles = n * (g.mcp / 4) * 17, nt = threads_of(n);
        return tiles <= nt ? (nt / tiles < 3 ? nt / tiles : 3) : 1;
    };
    auto smem_of = [&](int n) {
        const size_t rows = (size_t)n * R;
        const int tiles = n * (g.mcp / 4) * 17;
        const int T = split_of(n);
        return (size_t)4 * (8 * kE * kE + 4 * kE + (size_t)n * g.per_sz + 2 * rows * XS + 2 * rows * g.mcp + (size_t)(T - 1) * tiles * 16) + 16;
    };
    int gpc = 224 / R;
    if (gpc < 1) gpc = 1;
    if (gpc > 4) gpc = 4;
    while (gpc > 1 && smem_of(gpc) > 216 * 1024) --gpc;
    // two co-resident CTAs hide each other's barriers and fill a small batch in fewer waves: prefer the largest group that still
    // leaves room for two CTAs per SM
    for (int n = gpc; n >= 1; --n)
        if (smem_of(n) <= 110 * 1024) {
            gpc = n;
            break;
        }
    g.gpc = gpc;
    g.nthreads = threads_of(gpc);
    g.T = split_of(gpc);
    g.smem = smem_of(gpc);
    return g;
}

__device__ __forceinline__ void load_weights(const AttnArgs& a, float* WkvT, float* Wq, float* WoT, float* bin, float* bo) {
    const int tid = threadIdx.x, nt = blockDim.x;
    // WkvT[k][f] (f < 2E) = Win[E + f][k]  (k-major: the K / V projections walk k)
    for (int i = tid; i < 2 * kE * kE; i += nt) {
        const int f = i / kE, k = i - f * kE;
        WkvT[k * 2 * kE + f] = a.Win[kE * kE + i];
    }
    for (int i = tid; i < kE * kE; i += nt) {
        Wq[i] = a.Win[i];                                  // row-major [f][e]
        const int f = i / kE, k = i - f * kE;
        WoT[k * kE + f] = a.Wo[i];                         // WoT[k][f] = Wo[f][k]
    }
    for (int i = tid; i < 3 * kE; i += nt) bin[i] = a.bin[i];
    for (int i = tid; i < kE; i += nt) bo[i] = a.bo[i];
}

// K, V, K', V', c of the local graphs [0, ng)
__device__ __forceinline__ void graph_tables(const AttnArgs& a, int b0, int ng, float* per, int per_sz, const float* WkvT, const float* Wq,
                                             const float* WoT, const float* bin) {
    const int tid = threadIdx.x, nt = blockDim.x, M = a.M, H = a.heads, hd = kE / H, HM = H * M;
    const float scale = rsqrtf((float)hd);
    for (int i = tid; i < ng * M * kE; i += nt) {
        const int gl = i / (M * kE), r = i - gl * M * kE;
        per[gl * per_sz + r] = a.a[((int64_t)(b0 + gl) * M) * kE + r];
    }
    __syncthreads();
    // K = A Wk^T + bk, V = A Wv^T + bv : thread = (graph, token j, 4 output features of [K | V])
    for (int idx = tid; idx < ng * M * 16; idx += nt) {
        const int gl = idx / (M * 16), r = idx - gl * M * 16, j = r >> 4, c = r & 15;     // c < 8: K quad, else V quad
        const float* arow = per + gl * per_sz + j * kE;
        float4 acc = ld4s(bin + kE + 4 * c);
#pragma unroll 8
        for (int k = 0; k < kE; ++k) fma4(arow[k], ld4s(WkvT + k * 2 * kE + 4 * c), acc);
        float* dstp = per + gl * per_sz + M * kE + (c < 8 ? j * kE + 4 * c : M * kE + j * kE + 4 * (c - 8));
        st4s(dstp, acc);
    }
    __syncthreads();
    // K'_h[j][e] = scale sum_d K[j][h hd + d] Wq[h hd + d][e] ; V'_h[j][f] = sum_d V[j][h hd + d] WoT[h hd + d][f]
    for (int idx = tid; idx < ng * HM * 16; idx += nt) {
        const int gl = idx / (HM * 16), r = idx - gl * HM * 16, hj = r >> 4, c = r & 15, h = hj / M, j = hj - h * M;
        const float* base = per + gl * per_sz;
        const bool isK = c < 8;
        const float* srow = base + M * kE + (isK ? 0 : M * kE) + j * kE + h * hd;
        const float* wmat = (isK ? Wq : WoT) + h * hd * kE + 4 * (c & 7);
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int d = 0; d < hd; ++d) fma4(srow[d], ld4s(wmat + d * kE), acc);
        if (isK) {
            acc.x *= scale; acc.y *= scale; acc.z *= scale; acc.w *= scale;
        }
        float* dstp = per + gl * per_sz + 3 * M * kE + (isK ? 0 : HM * kE) + hj * kE + 4 * (c & 7);
        st4s(dstp, acc);
    }
    for (int idx = tid; idx < ng * HM; idx += nt) {
        const int gl = idx / HM, hj = idx - gl * HM, h = hj / M, j = hj - h * M;
        const float* krow = per + gl * per_sz + M * kE + j * kE + h * hd;
        float v = 0.f;
        for (int d = 0; d < hd; ++d) v = fmaf(krow[d], bin[h * hd + d], v);
        per[gl * per_sz + 3 * M * kE + 2 * HM * kE + hj] = v * scale;
    }
    __syncthreads();
}

// scores of one head for one row -> probabilities in s[] (s[j] = 0 for j >= M)
template <int MC>
__device__ __forceinline__ void row_softmax(const float4 (&x)[8], const float* Kp, const float* cb, int M, float (&s)[MC]) {
    float mx = -INFINITY;
#pragma unroll
    for (int j = 0; j < MC; ++j) {
        if (j < M) {
            const float* kr = Kp + j * kE;
            float d = cb[j];
#pragma unroll
            for (int k = 0; k < 8; ++k) d = dot4s(x[k], ld4s(kr + 4 * k), d);
            s[j] = d;
            mx = fmaxf(mx, d);
        } else {
            s[j] = -INFINITY;
        }
    }
    float den = 0.f;
#pragma unroll
    for (int j = 0; j < MC; ++j) {
        const float e = (j < M) ? __expf(s[j] - mx) : 0.f;
        s[j] = e;
        den += e;
    }
    const float inv = 1.f / den;
#pragma unroll
    for (int j = 0; j < MC; ++j) s[j] *= inv;
}

template <int MC>
__global__ void __launch_bounds__(288, 2) attn_rows_fwd_kernel(AttnArgs a, Geo geo) {
    IGCN_PDL_SYNC();
    extern __shared__ __align__(16) float smf[];
    const int tid = threadIdx.x, nt = blockDim.x, R = a.R, M = a.M, H = a.heads, HM = H * M;
    float* WkvT = smf;
    float* Wq = WkvT + 2 * kE * kE;
    float* WoT = Wq + kE * kE;
    float* bin = WoT + kE * kE;
    float* bo = bin + 3 * kE;
    float* per = bo + kE;                                   // offset 4 * (4 E E + 4 E): 16-byte aligned
    float* Xs = per + geo.gpc * geo.per_sz;
    load_weights(a, WkvT, Wq, WoT, bin, bo);
    __syncthreads();
    for (int b0 = blockIdx.x * geo.gpc; b0 < a.B; b0 += gridDim.x * geo.gpc) {
        const int ng = min(geo.gpc, a.B - b0), rows = ng * R;
        const float* xg = a.x + (int64_t)b0 * R * kE;
        for (int idx = tid; idx < rows * 8; idx += nt) {    // coalesced 16-byte loads -> padded rows
            const int row = idx >> 3, ch = idx & 7;
            st4s(Xs + row * XS + 4 * ch, ld4s(xg + (int64_t)idx * 4));
        }
        graph_tables(a, b0, ng, per, geo.per_sz, WkvT, Wq, WoT, bin);       // ends with a barrier (covers Xs too)
        for (int row = tid; row < rows; row += nt) {
            const int gl = row / R;
            const float* Kp = per + gl * geo.per_sz + 3 * M * kE;
            const float* Vp = Kp + HM * kE;
            const float* cb = Vp + HM * kE;
            float4 x[8], y[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                x[k] = ld4s(Xs + row * XS + 4 * k);
                y[k] = ld4s(bo + 4 * k);
            }
#pragma unroll 1
            for (int h = 0; h < H; ++h) {
                float s[MC];
                row_softmax<MC>(x, Kp + h * M * kE, cb + h * M, M, s);
#pragma unroll
                for (int j = 0; j < MC; ++j) {
                    if (j < M) {
                        const float* vr = Vp + (h * M + j) * kE;
#pragma unroll
                        for (int k = 0; k < 8; ++k) fma4(s[j], ld4s(vr + 4 * k), y[k]);
                    }
                }
            }
            float* yg = a.y + ((int64_t)b0 * R + row) * kE;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                float4 v = y[k];
                if (a.relu) {
                    v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f);
                }
                *reinterpret_cast<float4*>(yg + 4 * k) = v;
            }
        }
        __syncthreads();
    }
}

// ---- backward -------------------------------------------------------------------------------------------------------------
// acc layout (= AttnArgs::partials row, shared with the staged kernel): [dWin (3E,E) | dbin (3E) | dWo (E,E) | dbo (E)]
template <int MC>
__global__ void __launch_bounds__(288) attn_rows_bwd_kernel(AttnArgs a, Geo geo) {
    IGCN_PDL_SYNC();
    extern __shared__ __align__(16) float smf[];
    const int tid = threadIdx.x, nt = blockDim.x, R = a.R, M = a.M, H = a.heads, hd = kE / H, HM = H * M, MCP = geo.mcp, JQ = MCP / 4;
    const int HMP = (HM + 3) & ~3;
    const float scale = rsqrtf((float)hd);
    float* WkvT = smf;
    float* Wq = WkvT + 2 * kE * kE;
    float* WoT = Wq + kE * kE;
    float* bin = WoT + kE * kE;
    float* bo = bin + 3 * kE;
    float* WqT = bo + kE;                                   // extra layouts of the backward chain (lanes walk the contiguous dimension)
    float* WoO = WqT + kE * kE;
    float* WkvO = WoO + kE * kE;
    float* per = WkvO + 2 * kE * kE;
    const int rows_max = geo.gpc * R;
    float* Xs = per + geo.gpc * geo.per_sz;                 // rows x XS, column 32 = 1
    float* Ys = Xs + rows_max * XS;                         // masked dY rows x XS
    float* Pst = Ys + rows_max * XS;                        // rows x MCP : P of the current head, column M = 1
    float* Dst = Pst + rows_max * MCP;                      // rows x MCP : dS of the current head
    float* scratch = Dst + rows_max * MCP;                  // 3 x tiles x 16 partial tiles
    // per-graph table offsets
    const int oK = M * kE, oV = 2 * M * kE, oKp = 3 * M * kE, oVp = oKp + HM * kE, oC = oVp + HM * kE, oDC = oC + HMP;
    const int oDKp = oDC + HMP, oDVp = oDKp + H * MCP * XS, oDK = oDVp + H * MCP * XS, oDV = oDK + M * kE;
    float* accg = a.partials + (int64_t)blockIdx.x * a.P;
    const int oBin = 3 * kE * kE, oWo = oBin + 3 * kE, oBo = oWo + kE * kE;
    for (int i = tid; i < a.P; i += nt) accg[i] = 0.f;
    load_weights(a, WkvT, Wq, WoT, bin, bo);
    for (int i = tid; i < kE * kE; i += nt) {
        const int f = i / kE, e = i - f * kE;
        WqT[e * kE + f] = a.Win[i];                        // WqT[e][f] = Wq[f][e]
        WoO[i] = a.Wo[i];                                  // row-major [f'][f]
    }
    for (int i = tid; i < 2 * kE * kE; i += nt) WkvO[i] = a.Win[kE * kE + i];     // row-major [Wk ; Wv]
    __syncthreads();
    const int tiles_g = JQ * 17;                            // per graph: JQ x 8 tiles of dV', JQ x 9 tiles of dK'
    for (int b0 = blockIdx.x * geo.gpc; b0 < a.B; b0 += gridDim.x * geo.gpc) {
        const int ng = min(geo.gpc, a.B - b0), rows = ng * R;
        const float* xg = a.x + (int64_t)b0 * R * kE;
        const float* gg = a.gy + (int64_t)b0 * R * kE;
        const float* yg = a.yout + (int64_t)b0 * R * kE;
        for (int idx = tid; idx < rows * 8; idx += nt) {
            const int row = idx >> 3, ch = idx & 7;
            st4s(Xs + row * XS + 4 * ch, ld4s(xg + (int64_t)idx * 4));
            float4 g = ld4s(gg + (int64_t)idx * 4);
            if (a.relu) {
                const float4 yv = ld4s(yg + (int64_t)idx * 4);
                if (!(yv.x > 0.f)) g.x = 0.f;
                if (!(yv.y > 0.f)) g.y = 0.f;
                if (!(yv.z > 0.f)) g.z = 0.f;
                if (!(yv.w > 0.f)) g.w = 0.f;
            }
            st4s(Ys + row * XS + 4 * ch, g);
        }
        for (int row = tid; row < rows; row += nt) {
            st4s(Xs + row * XS + 32, make_float4(1.f, 0.f, 0.f, 0.f));
            for (int j = MC; j < MCP; ++j) {                // stage columns beyond the register arrays (P: ones column at j = M)
                Pst[row * MCP + j] = (j == M) ? 1.f : 0.f;
                Dst[row * MCP + j] = 0.f;
            }
        }
        for (int i = tid; i < ng * 2 * H * MCP * XS; i += nt) {       // dK', dV' accumulators
            const int gl = i / (2 * H * MCP * XS), r = i - gl * (2 * H * MCP * XS);
            per[gl * geo.per_sz + oDKp + r] = 0.f;
        }
        graph_tables(a, b0, ng, per, geo.per_sz, WkvT, Wq, WoT, bin);
        // ---- row pass, one head at a time; the reductions over rows follow each head ---------------------------------------
        const bool active = tid < rows;                     // rows <= nthreads by construction (checked on the host)
        const int row = tid, gl = active ? row / R : 0;
        const float* tb = per + gl * geo.per_sz;
        float4 x[8], dy[8], dx[8];
        if (active) {
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                x[k] = ld4s(Xs + row * XS + 4 * k);
                dy[k] = ld4s(Ys + row * XS + 4 * k);
                dx[k] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
#pragma unroll 1
        for (int h = 0; h < H; ++h) {
            if (active) {
                float s[MC], dp[MC];
                row_softmax<MC>(x, tb + oKp + h * M * kE, tb + oC + h * M, M, s);
                float rowdot = 0.f;
#pragma unroll
                for (int j = 0; j < MC; ++j) {
                    if (j < M) {
                        const float* vr = tb + oVp + (h * M + j) * kE;
                        float d = 0.f;
#pragma unroll
                        for (int k = 0; k < 8; ++k) d = dot4s(dy[k], ld4s(vr + 4 * k), d);
                        dp[j] = d;
                        rowdot = fmaf(s[j], d, rowdot);
                    } else {
                        dp[j] = 0.f;
                    }
                }
#pragma unroll
                for (int j = 0; j < MC; ++j) {
                    float ds = 0.f;
                    if (j < M) {
                        ds = s[j] * (dp[j] - rowdot);
                        const float* kr = tb + oKp + (h * M + j) * kE;
#pragma unroll
                        for (int k = 0; k < 8; ++k) fma4(ds, ld4s(kr + 4 * k), dx[k]);
                    }
                    dp[j] = ds;
                    if (j == M) s[j] = 1.f;                 // ones column of P (column sums of dY come out of the dV' reduction)
                }
                // stage rows as 16-byte stores: stride MCP floats (a multiple of 4, 80 B at M = 19) is conflict free for STS.128
#pragma unroll
                for (int q = 0; q < MC / 4; ++q) {
                    if (4 * q < MCP) {
                        st4s(Pst + row * MCP + 4 * q, make_float4(s[4 * q], s[4 * q + 1], s[4 * q + 2], s[4 * q + 3]));
                        st4s(Dst + row * MCP + 4 * q, make_float4(dp[4 * q], dp[4 * q + 1], dp[4 * q + 2], dp[4 * q + 3]));
                    }
                }
            }
            __syncthreads();
            // reductions over the rows of each graph: dV'_h (MCP x 32) = P^T dY ; dK'_h (MCP x 36) = dS^T [X | 1]
            // tile = 4 (j) x 4 (columns).  When there are fewer tiles than threads, the rows of a graph are split over T thread
            // groups and the partial tiles are added in group order; otherwise a thread walks several tiles.
            {
                const int tiles = ng * tiles_g;
                const int T = geo.T;                        // sized for a full pass (a short last pass keeps the same split)
                auto tile_sum = [&](int tl, int t, float (&acc)[4][4], int& gl2, int& jq, int& cq, bool& isV) {
                    gl2 = tl / tiles_g;
                    const int r = tl - gl2 * tiles_g;
                    isV = r < JQ * 8;
                    if (isV) { jq = r >> 3; cq = r & 7; } else { const int r2 = r - JQ * 8; jq = r2 / 9; cq = r2 - jq * 9; }
                    const float* lhs = (isV ? Pst : Dst) + (gl2 * R) * MCP + 4 * jq;
                    const float* rhs = (isV ? Ys : Xs) + (gl2 * R) * XS + 4 * cq;
                    const int i0 = (R * t) / T, i1 = (R * (t + 1)) / T;
#pragma unroll
                    for (int p = 0; p < 4; ++p)
#pragma unroll
                        for (int q = 0; q < 4; ++q) acc[p][q] = 0.f;
#pragma unroll 4
                    for (int i = i0; i < i1; ++i) {
                        const float4 l = ld4s(lhs + i * MCP), r4 = ld4s(rhs + i * XS);
                        const float lv[4] = {l.x, l.y, l.z, l.w}, rv[4] = {r4.x, r4.y, r4.z, r4.w};
#pragma unroll
                        for (int p = 0; p < 4; ++p)
#pragma unroll
                            for (int q = 0; q < 4; ++q) acc[p][q] = fmaf(lv[p], rv[q], acc[p][q]);
                    }
                };
                auto tile_store = [&](const float (&acc)[4][4], int gl2, int jq, int cq, bool isV) {
                    float* dstp = per + gl2 * geo.per_sz + (isV ? oDVp : oDKp) + (h * MCP + 4 * jq) * XS + 4 * cq;
#pragma unroll
                    for (int p = 0; p < 4; ++p) st4s(dstp + p * XS, make_float4(acc[p][0], acc[p][1], acc[p][2], acc[p][3]));
                };
                float acc[4][4];
                int gl2 = 0, jq = 0, cq = 0;
                bool isV = false;
                if (T == 1) {
                    for (int tl = tid; tl < tiles; tl += nt) {
                        tile_sum(tl, 0, acc, gl2, jq, cq, isV);
                        tile_store(acc, gl2, jq, cq, isV);
                    }
                } else {
                    const int t = tid / tiles, tl = tid - t * tiles;
                    const bool worker = t < T;
                    if (worker) {
                        tile_sum(tl, t, acc, gl2, jq, cq, isV);
                        if (t > 0) {
                            float* sc = scratch + ((t - 1) * tiles + tl) * 16;
#pragma unroll
                            for (int p = 0; p < 4; ++p) st4s(sc + 4 * p, make_float4(acc[p][0], acc[p][1], acc[p][2], acc[p][3]));
                        }
                    }
                    __syncthreads();
                    if (worker && t == 0) {
                        for (int u = 1; u < T; ++u) {
                            const float* sc = scratch + ((u - 1) * tiles + tl) * 16;
#pragma unroll
                            for (int p = 0; p < 4; ++p) {
                                const float4 v = ld4s(sc + 4 * p);
                                acc[p][0] += v.x; acc[p][1] += v.y; acc[p][2] += v.z; acc[p][3] += v.w;
                            }
                        }
                        tile_store(acc, gl2, jq, cq, isV);
                    }
                }
                __syncthreads();
            }
        }
        if (active) {
            float* dxg = a.dx + ((int64_t)b0 * R + row) * kE;
#pragma unroll
            for (int k = 0; k < 8; ++k) *reinterpret_cast<float4*>(dxg + 4 * k) = dx[k];
        }
        // ---- chain back to K, V, the tokens and the parameters (M x 32 sized) ----------------------------------------------
        // dK[j][f] = scale (<dK'_h[j], Wq[f]> + dc_h[j] bq[f]) ; dV[j][f] = <dV'_h[j], WoT[f]>     (f = h hd + d)
        for (int idx = tid; idx < ng * M * 16; idx += nt) {
            const int gl3 = idx / (M * 16), r = idx - gl3 * M * 16, j = r >> 4, c = r & 15, fq = c & 7;
            const bool isK = c < 8;
            const int h = (4 * fq) / hd;
            float* base = per + gl3 * geo.per_sz;
            const float* src = base + (isK ? oDKp : oDVp) + (h * MCP + j) * XS;      // dK'_h[j][:] / dV'_h[j][:]
            const float* wm = (isK ? WqT : WoO) + 4 * fq;                             // [e][4fq..] / [f'][4fq..]
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 8
            for (int e = 0; e < kE; ++e) fma4(src[e], ld4s(wm + e * kE), v);
            if (isK) {
                const float dc = src[32];
                const float4 bq = ld4s(bin + 4 * fq);
                v.x = scale * (v.x + dc * bq.x); v.y = scale * (v.y + dc * bq.y);
                v.z = scale * (v.z + dc * bq.z); v.w = scale * (v.w + dc * bq.w);
            }
            st4s(base + (isK ? oDK : oDV) + j * kE + 4 * fq, v);
        }
        // dWq[f][4q..] += scale sum_j K[j][f] dK'_h[j][4q..] ; dWo[f'][4q..] += sum_j dV'_h[j][f'] V[j][4q..]   (h = head of the f side)
        // dbq[f] += scale sum_j K[j][f] dc_h[j] ; dbo[f] += column sums of dY (the ones row of P, head 0)
        for (int idx = tid; idx < 2 * kE * 8 + kE; idx += nt) {
            if (idx < 2 * kE * 8) {
                const bool isQ = idx < kE * 8;
                const int i2 = isQ ? idx : idx - kE * 8, f = i2 >> 3, q = i2 & 7;
                const int h = isQ ? f / hd : (4 * q) / hd;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                for (int gl3 = 0; gl3 < ng; ++gl3) {
                    const float* base = per + gl3 * geo.per_sz;
                    const float* sc1 = isQ ? base + oK + f : base + oDVp + h * MCP * XS + f;           // scalar operand, walks j
                    const int ld1 = isQ ? kE : XS;
                    const float* vec = isQ ? base + oDKp + h * MCP * XS + 4 * q : base + oV + 4 * q;  // float4 operand, walks j
                    const int ld2 = isQ ? XS : kE;
#pragma unroll 4
                    for (int j = 0; j < M; ++j) fma4(sc1[j * ld1], ld4s(vec + j * ld2), v);
                }
                float* d = accg + (isQ ? 0 : oWo) + f * kE + 4 * q;
                const float sf = isQ ? scale : 1.f;
                d[0] += sf * v.x; d[1] += sf * v.y; d[2] += sf * v.z; d[3] += sf * v.w;
            } else {
                const int f = idx - 2 * kE * 8, h = f / hd;
                float v = 0.f, vb = 0.f;
                for (int gl3 = 0; gl3 < ng; ++gl3) {
                    const float* base = per + gl3 * geo.per_sz;
#pragma unroll 4
                    for (int j = 0; j < M; ++j) v = fmaf(base[oK + j * kE + f], base[oDKp + (h * MCP + j) * XS + 32], v);
                    vb += base[oDVp + M * XS + f];
                }
                accg[oBin + f] += scale * v;
                accg[oBo + f] += vb;
            }
        }
        __syncthreads();
        // token gradient dA[j][k] = <dK[j], Wk[:,k]> + <dV[j], Wv[:,k]> ; dWk/dWv[f][k] += sum_j d{K,V}[j][f] A[j][k] ; dbk/dbv
        for (int idx = tid; idx < ng * M * 8; idx += nt) {
            const int gl3 = idx / (M * 8), r = idx - gl3 * M * 8, j = r >> 3, kq = r & 7;
            const float* base = per + gl3 * geo.per_sz;
            const float* dk = base + oDK + j * kE;
            const float* dv = base + oDV + j * kE;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 8
            for (int f = 0; f < kE; ++f) {
                fma4(dk[f], ld4s(WkvO + f * kE + 4 * kq), v);
                fma4(dv[f], ld4s(WkvO + (kE + f) * kE + 4 * kq), v);
            }
            *reinterpret_cast<float4*>(a.da + ((int64_t)(b0 + gl3) * M + j) * kE + 4 * kq) = v;
        }
        for (int idx = tid; idx < 2 * kE * 8 + 2 * kE; idx += nt) {
            if (idx < 2 * kE * 8) {                          // dW{k,v}[f2][4q..] += sum_j d{K,V}[j][f2] A[j][4q..]
                const int f2 = idx >> 3, q = idx & 7;
                const int o = (f2 < kE) ? oDK + f2 : oDV + (f2 - kE);
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                for (int gl3 = 0; gl3 < ng; ++gl3) {
                    const float* base = per + gl3 * geo.per_sz;
#pragma unroll 4
                    for (int j = 0; j < M; ++j) fma4(base[o + j * kE], ld4s(base + j * kE + 4 * q), v);
                }
                float* d = accg + kE * kE + f2 * kE + 4 * q;
                d[0] += v.x; d[1] += v.y; d[2] += v.z; d[3] += v.w;
            } else {
                const int f2 = idx - 2 * kE * 8;
                const int o = (f2 < kE) ? oDK + f2 : oDV + (f2 - kE);
                float v = 0.f;
                for (int gl3 = 0; gl3 < ng; ++gl3) {
                    const float* base = per + gl3 * geo.per_sz;
                    for (int j = 0; j < M; ++j) v += base[o + j * kE];
                }
                accg[oBin + kE + f2] += v;
            }
        }
        __syncthreads();
    }
}

}  // namespace rows
}  // namespace igcn

// Dense products of the path on the 5th-generation tensor cores (tcgen05 + TMEM + TMA), fp32-accurate.
//
// Which products: the imaging-SNP fusion heads lin1 / lin1_regr (kernel/sgcn_img_snp.py:286-305; K = R*L*H+32 (+R*F0) =
// 2 912 .. 9 272, N = 64, M = batch) forward and both backward products, and the (B x B)(B x D) Laplacian product of
// consist_loss (kernel/sgcn_img_snp.py:183-196).  profiles/r1_bench_config4.json: on FFMA tiles these are 70 % of the
// config-4 step -- the "genuine dense contraction" the north star asks to see before using tensor cores.
//
// Precision: one TF32 product (10-bit mantissa) cannot hold the 1e-4 parity bar through the cancellation in the Laplacian
// form, so every operand is split as x = hi + lo (hi = the top 19 bits of x, lo = x - hi: exact) and C = Ah*Bh + Ah*Bl + Al*Bh is
// accumulated in fp32 in TMEM ("3xTF32"); the dropped Al*Bl term is 2^-22 relative.  Measured against fp64: ~3e-7.
//
// Shape of the kernel (one CTA per 128 x BN output tile and K split; 192 threads):
//   warp 4   : TMA producer -- per 32-wide k block four cp.async.bulk.tensor loads (A_hi, A_lo, B_hi, B_lo; SWIZZLE_128B,
//              out-of-bounds rows / k zero-filled by the tensor map) into a 3-4 stage shared-memory ring, mbarrier expect_tx
//   warp 5   : allocates TMEM (BN columns x 128 lanes fp32), one elected lane issues 12 tcgen05.mma.kind::tf32
//              (128 x BN x 8) per k block from shared-memory descriptors; tcgen05.commit releases the ring slot
//   warps 0-3: epilogue -- tcgen05.ld 32 lanes x 32 columns per warp, bias/ReLU, stores straight to the (up to three)
//              destination segments, or to the split-K partial buffer (summed in a fixed order by tc_reduce_kernel)
// Operands are K-major: C[m][n] = sum_k A[m][k] B[n][k]; transposed operands are produced by the split pass (tc_split_kernel),
// which is also where cat(...) and the ReLU mask of the backward are folded in.
#include <cuda.h>

#include "common.cuh"

namespace igcn {
namespace tc {

constexpr int BM = 128, BK = 32;            // BK fp32 = 128 bytes = one SWIZZLE_128B row
constexpr int THREADS = 192;
constexpr int A_TILE = BM * BK * 4;         // 16 KB
constexpr long long SPIN_LIMIT = 4000000000LL;   // ~2 s of SM clocks: a protocol bug traps instead of hanging the GPU

template <int BN> struct Cfg {
    static constexpr int STAGES = BN == 64 ? 4 : 3;
    static constexpr int B_TILE = BN * BK * 4;
    static constexpr int STAGE_BYTES = 2 * A_TILE + 2 * B_TILE;
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 + 256;
};

struct SegDst {
    float* p[3];
    int w[3];
    int ld[3];
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, int who) {
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > SPIN_LIMIT) {
            printf("igcn tc_gemm: mbarrier wait timed out (role %d, block %d,%d,%d)\n", who, blockIdx.x, blockIdx.y, blockIdx.z);
            __trap();
        }
    }
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int c_inner, int c_outer, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
                 "l"(map), "r"(bar), "r"(c_inner), "r"(c_outer)
                 : "memory");
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// shared-memory matrix descriptor: K-major tile, rows of 128 B, SWIZZLE_128B, 8-row groups 1024 B apart (cute::UMMA::SmemDescriptor)
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr) {
    uint64_t d = 0;
    d |= (uint64_t)((addr >> 4) & 0x3FFF);          // start address, 16-byte units
    d |= (uint64_t)1 << 16;                         // leading byte offset (unused for swizzled K-major), 16-byte units
    d |= (uint64_t)(1024 >> 4) << 32;               // stride byte offset between 8-row groups
    d |= (uint64_t)1 << 46;                         // descriptor version (Blackwell)
    d |= (uint64_t)2 << 61;                         // SWIZZLE_128B
    return d;
}

template <int BN>
__global__ void __launch_bounds__(THREADS, 1)
tc_gemm_kernel(const __grid_constant__ CUtensorMap tm_ah, const __grid_constant__ CUtensorMap tm_al,
               const __grid_constant__ CUtensorMap tm_bh, const __grid_constant__ CUtensorMap tm_bl, int M, int N, int kblocks,
               int kb_per_split, const float* __restrict__ bias, int relu, SegDst dst, float* __restrict__ partials) {
    IGCN_PDL_SYNC();
    using C = Cfg<BN>;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;            // SWIZZLE_128B tiles need 1024-byte alignment
    const uint32_t bars = base + C::STAGES * C::STAGE_BYTES;                 // full[STAGES], empty[STAGES], tmem_full, tmem_ptr
    const uint32_t bar_full = bars, bar_empty = bars + 8 * C::STAGES, bar_tmem = bars + 16 * C::STAGES;
    const uint32_t tmem_slot = bars + 16 * C::STAGES + 8;
    volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN, split = blockIdx.z;
    const int kb0 = split * kb_per_split;
    const int nkb = min(kblocks - kb0, kb_per_split);                        // >= 1 (host guarantees)

    if (threadIdx.x == 0) {
        for (int s = 0; s < C::STAGES; ++s) {
            mbar_init(bar_full + 8 * s, 1);
            mbar_init(bar_empty + 8 * s, 1);
        }
        mbar_init(bar_tmem, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 5) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "n"(BN) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_d = *tmem_slot_ptr;

    if (warp == 4) {
        if (lane == 0) {
            asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_ah) : "memory");
            asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_al) : "memory");
            asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_bh) : "memory");
            asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_bl) : "memory");
            for (int i = 0; i < nkb; ++i) {
                const int s = i % C::STAGES;
                const uint32_t ph = (uint32_t)(i / C::STAGES) & 1u;
                mbar_wait(bar_empty + 8 * s, ph ^ 1u, 0);
                const uint32_t sa = base + s * C::STAGE_BYTES;
                const uint32_t full = bar_full + 8 * s;
                mbar_expect_tx(full, (uint32_t)C::STAGE_BYTES);
                const int k = (kb0 + i) * BK;
                tma_load_2d(sa, &tm_ah, k, m0, full);
                tma_load_2d(sa + A_TILE, &tm_al, k, m0, full);
                tma_load_2d(sa + 2 * A_TILE, &tm_bh, k, n0, full);
                tma_load_2d(sa + 2 * A_TILE + C::B_TILE, &tm_bl, k, n0, full);
            }
        }
    } else if (warp == 5) {
        if (lane == 0) {
            // instruction descriptor (cute::UMMA::InstrDescriptor): D = f32, A = B = tf32, both K-major, N >> 3, M >> 4
            const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
            for (int i = 0; i < nkb; ++i) {
                const int s = i % C::STAGES;
                const uint32_t ph = (uint32_t)(i / C::STAGES) & 1u;
                mbar_wait(bar_full + 8 * s, ph, 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t sa = base + s * C::STAGE_BYTES;
                const uint64_t ah = smem_desc(sa), al = smem_desc(sa + A_TILE);
                const uint64_t bh = smem_desc(sa + 2 * A_TILE), bl = smem_desc(sa + 2 * A_TILE + C::B_TILE);
#pragma unroll
                for (int kk = 0; kk < BK / 8; ++kk) {                 // 8 tf32 = 32 bytes per MMA along K: +2 in 16-byte units
                    const uint64_t o = (uint64_t)(kk * 2);
                    umma_tf32(tmem_d, al + o, bh + o, idesc, (i | kk) != 0);      // small terms first
                    umma_tf32(tmem_d, ah + o, bl + o, idesc, 1u);
                    umma_tf32(tmem_d, ah + o, bh + o, idesc, 1u);
                }
                umma_commit(bar_empty + 8 * s);                        // frees the ring slot when these MMAs have read it
            }
            umma_commit(bar_tmem);                                     // accumulator complete
        }
    } else {
        // ---- epilogue: warp w owns TMEM lanes [32w, 32w+32) = output rows m0 + 32w + lane -------------------------------
        mbar_wait(bar_tmem, 0u, 2);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        // The accumulator arrives with lane = output row; stored that way every lane writes its own row (32 sectors per store
        // instruction, 4 bytes each: the 8x write amplification ncu showed on the Laplacian product).  Each warp turns its 32 x 32
        // block around through a padded tile in the (now idle) operand ring, so a store instruction covers 128 contiguous bytes
        // of ONE output row.
        float* tb = reinterpret_cast<float*>(smem_raw + (base - smem_u32(smem_raw))) + warp * (32 * 33);
        const int mrow0 = m0 + warp * 32;
#pragma unroll 1
        for (int c = 0; c < BN; c += 32) {
            __syncwarp();                                             // converged for the .aligned TMEM load; previous tile readers done
            uint32_t r[32];
            const uint32_t taddr = tmem_d + ((uint32_t)(warp * 32) << 16) + (uint32_t)c;
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                  "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
                  "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
                  "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                : "r"(taddr));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int j = 0; j < 32; ++j) tb[lane * 33 + j] = __uint_as_float(r[j]);
            __syncwarp();
            int n = n0 + c + lane;                                    // this lane's output column for the whole block
            if (n >= N) continue;
            if (partials) {                                           // split-K partial tile, raw
                float* pcol = partials + (int64_t)split * M * N + n;
#pragma unroll 4
                for (int rr = 0; rr < 32; ++rr)
                    if (mrow0 + rr < M) pcol[(int64_t)(mrow0 + rr) * N] = tb[rr * 33 + lane];
                continue;
            }
            const float bv = bias ? bias[n] : 0.f;
            int sidx = 0;
            if (n >= dst.w[0]) {
                n -= dst.w[0];
                sidx = 1;
                if (n >= dst.w[1]) {
                    n -= dst.w[1];
                    sidx = 2;
                }
            }
            float* p = dst.p[sidx];
            if (!p) continue;
            const int64_t ld = dst.ld[sidx];
#pragma unroll 4
            for (int rr = 0; rr < 32; ++rr) {
                if (mrow0 + rr >= M) break;
                float v = tb[rr * 33 + lane] + bv;
                if (relu) v = fmaxf(v, 0.f);
                p[(int64_t)(mrow0 + rr) * ld + n] = v;
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 5) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "n"(BN) : "memory");
    }
}

// out segment[m][n] = act(sum_s partials[s][m][n] + bias[n]) in split order (deterministic)
__global__ void __launch_bounds__(256) tc_reduce_kernel(const float* __restrict__ partials, const float* __restrict__ bias, int M, int N,
                                                        int S, int relu, SegDst dst) {
    IGCN_PDL_SYNC();
    const int64_t idx = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (idx >= (int64_t)M * N) return;
    const int m = (int)(idx / N);
    int n = (int)(idx - (int64_t)m * N);
    float v = 0.f;
    for (int s = 0; s < S; ++s) v += partials[(int64_t)s * M * N + idx];
    if (bias) v += bias[n];
    if (relu) v = fmaxf(v, 0.f);
    int sidx = 0;
    if (n >= dst.w[0]) {
        n -= dst.w[0];
        sidx = 1;
        if (n >= dst.w[1]) {
            n -= dst.w[1];
            sidx = 2;
        }
    }
    float* p = dst.p[sidx];
    if (p) p[(int64_t)m * dst.ld[sidx] + n] = v;
}

// ---- operand preparation: x -> (hi, lo) tf32 pair, optional ReLU mask (x *= mask > 0), optional transpose ----------------
struct SplitJob {
    const float* src;     // (rows, cols), row stride ld_src; NULL = constant 1.0
    const float* mask;    // same geometry as src, or NULL
    const float* sub;     // per source column, or NULL: x <- x - sub[c] (the centring of the Laplacian product)
    float *hi, *lo;       // destination matrices, row stride ld_dst
    int rows, cols, ld_src, ld_dst, row_off, col_off, transpose;
};
constexpr int MAX_JOBS = 8;
struct SplitJobs {
    SplitJob j[MAX_JOBS];
};

// x = hi + lo exactly: hi keeps the 19 bits kind::tf32 reads (1 + 8 + 10), lo = x - hi is exact in fp32 and the tensor core reads
// its top 10 mantissa bits, so what is dropped is < 2^-21 |x| (cvt.rna.tf32 is emulated on sm_100a by a 4-instruction sequence per
// conversion; this is a LOP3 and an FADD -- the split of mma_util.cuh)
__device__ __forceinline__ void split_tf32(float x, float& hi, float& lo) {
    hi = __uint_as_float(__float_as_uint(x) & 0xffffe000u);
    lo = x - hi;
}

// grid: (tiles of 32 x 32 over the largest job, njobs); block 32 x 8
__global__ void __launch_bounds__(256) tc_split_kernel(SplitJobs jobs) {
    IGCN_PDL_SYNC();
    __shared__ float tile[32][33];
    const SplitJob& J = jobs.j[blockIdx.y];
    const int tiles_c = (J.cols + 31) / 32, tiles_r = (J.rows + 31) / 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int t = blockIdx.x; t < tiles_c * tiles_r; t += gridDim.x) {
        const int r0 = (t / tiles_c) * 32, c0 = (t - (t / tiles_c) * tiles_c) * 32;
        if (!J.transpose) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int r = r0 + ty + i * 8, c = c0 + tx;
                if (r < J.rows && c < J.cols) {
                    float x = J.src ? J.src[(int64_t)r * J.ld_src + c] : 1.f;
                    if (J.mask && !(J.mask[(int64_t)r * J.ld_src + c] > 0.f)) x = 0.f;
                    if (J.sub) x -= J.sub[c];
                    float h, l;
                    split_tf32(x, h, l);
                    const int64_t o = (int64_t)(J.row_off + r) * J.ld_dst + J.col_off + c;
                    J.hi[o] = h;
                    J.lo[o] = l;
                }
            }
        } else {
            __syncthreads();
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int r = r0 + ty + i * 8, c = c0 + tx;
                float x = 0.f;
                if (r < J.rows && c < J.cols) {
                    x = J.src ? J.src[(int64_t)r * J.ld_src + c] : 1.f;
                    if (J.mask && !(J.mask[(int64_t)r * J.ld_src + c] > 0.f)) x = 0.f;
                    if (J.sub) x -= J.sub[c];
                }
                tile[ty + i * 8][tx] = x;
            }
            __syncthreads();
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int c = c0 + ty + i * 8, r = r0 + tx;           // destination row = source column
                if (r < J.rows && c < J.cols) {
                    float h, l;
                    split_tf32(tile[tx][ty + i * 8], h, l);
                    const int64_t o = (int64_t)(J.row_off + c) * J.ld_dst + J.col_off + r;
                    J.hi[o] = h;
                    J.lo[o] = l;
                }
            }
        }
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

// (rows, K) fp32, K contiguous, row pitch ld floats; box = 32 k x box_rows; out-of-bounds elements read as zero
static int make_map(CUtensorMap* map, const float* ptr, int64_t rows, int64_t K, int64_t ld, int box_rows) {
    EncodeTiledFn fn = encode_fn();
    IGCN_REQUIRE(fn, IGCN_ERR_UNSUPPORTED, "tc_gemm: cuTensorMapEncodeTiled is not available from this driver");
    cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
    cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    IGCN_REQUIRE(r == CUDA_SUCCESS, IGCN_ERR_BAD_ARG, "tc_gemm: cuTensorMapEncodeTiled failed (code %d; rows=%lld K=%lld ld=%lld)", (int)r,
                 (long long)rows, (long long)K, (long long)ld);
    return IGCN_OK;
}

template <int BN>
static int launch_gemm(const float* a_hi, const float* a_lo, int64_t lda, const float* b_hi, const float* b_lo, int64_t ldb, int64_t M,
                       int64_t N, int64_t K, const float* bias, int relu, const SegDst& dst, float* partials, int64_t S, cudaStream_t st) {
    using C = Cfg<BN>;
    CUtensorMap ta, tal, tb, tbl;
    int rc;
    if ((rc = make_map(&ta, a_hi, M, K, lda, BM))) return rc;
    if ((rc = make_map(&tal, a_lo, M, K, lda, BM))) return rc;
    if ((rc = make_map(&tb, b_hi, N, K, ldb, BN))) return rc;
    if ((rc = make_map(&tbl, b_lo, N, K, ldb, BN))) return rc;
    rc = allow_smem(tc_gemm_kernel<BN>, C::SMEM_BYTES, "tc_gemm");
    if (rc) return rc;
    const int kblocks = (int)((K + BK - 1) / BK);
    const int kb_per = (int)((kblocks + S - 1) / S);
    dim3 grid((unsigned)((M + BM - 1) / BM), (unsigned)((N + BN - 1) / BN), (unsigned)S);
    igcn::launch_k(tc_gemm_kernel<BN>, dim3(grid), dim3(THREADS), C::SMEM_BYTES, st, ta, tal, tb, tbl, (int)M, (int)N, kblocks, kb_per, S > 1 ? nullptr : bias,
                                                            S > 1 ? 0 : relu, dst, S > 1 ? partials : nullptr);
    IGCN_CHECK_LAUNCH("tc_gemm");
    if (S > 1) {
        igcn::launch_k(tc_reduce_kernel, dim3((unsigned)((M * N + 255) / 256)), dim3(256), 0, st, partials, bias, (int)M, (int)N, (int)S, relu, dst);
        IGCN_CHECK_LAUNCH("tc_reduce");
    }
    return IGCN_OK;
}

}  // namespace tc
}  // namespace igcn

using namespace igcn;

extern "C" int64_t igcn_tc_gemm_splits(int64_t M, int64_t N, int64_t K) {
    const int64_t bn = N <= 64 ? 64 : 128;
    const int64_t tiles = ((M + tc::BM - 1) / tc::BM) * ((N + bn - 1) / bn);
    const int64_t kblocks = (K + tc::BK - 1) / tc::BK;
    int64_t S = sm_count() / (tiles > 0 ? tiles : 1);
    if (S > kblocks / 2) S = kblocks / 2;                     // at least two k blocks per split
    if (S < 1) S = 1;
    // every split must own at least one k block: S <= ceil(kblocks / ceil(kblocks / S))
    while (S > 1 && (S - 1) * ((kblocks + S - 1) / S) >= kblocks) --S;
    return S;
}

extern "C" int igcn_tc_gemm(const float* a_hi, const float* a_lo, int64_t lda, const float* b_hi, const float* b_lo, int64_t ldb, int64_t M,
                            int64_t N, int64_t K, const float* bias, int64_t relu, float* d0, float* d1, float* d2,
                            const int64_t* host_dst_widths, const int64_t* host_dst_strides, float* partials, int64_t S, void* stream) {
    IGCN_REQUIRE(a_hi && a_lo && b_hi && b_lo && host_dst_widths && host_dst_strides, IGCN_ERR_BAD_ARG, "tc_gemm: null pointer");
    IGCN_REQUIRE(M > 0 && N > 0 && K > 0, IGCN_ERR_BAD_ARG, "tc_gemm: bad size");
    IGCN_REQUIRE(lda >= K && ldb >= K && (lda & 3) == 0 && (ldb & 3) == 0, IGCN_ERR_BAD_ARG,
                 "tc_gemm: operand row pitches must be >= K and multiples of 4 floats (TMA: 16-byte pitch)");
    IGCN_REQUIRE(((((uintptr_t)a_hi) | ((uintptr_t)a_lo) | ((uintptr_t)b_hi) | ((uintptr_t)b_lo)) & 15) == 0, IGCN_ERR_BAD_ARG,
                 "tc_gemm: operands must be 16-byte aligned");
    IGCN_REQUIRE(S == igcn_tc_gemm_splits(M, N, K), IGCN_ERR_BAD_ARG, "tc_gemm: S must be igcn_tc_gemm_splits()");
    IGCN_REQUIRE(S == 1 || partials, IGCN_ERR_BAD_ARG, "tc_gemm: split-K needs the partials workspace (S*M*N floats)");
    IGCN_REQUIRE(M < (1ll << 31) && N < (1ll << 31) && K < (1ll << 31), IGCN_ERR_UNSUPPORTED, "tc_gemm: dimension too large");
    tc::SegDst dst;
    float* ds[3] = {d0, d1, d2};
    int64_t wsum = 0;
    for (int i = 0; i < 3; ++i) {
        IGCN_REQUIRE(host_dst_widths[i] >= 0 && (host_dst_widths[i] == 0 || host_dst_strides[i] >= host_dst_widths[i]), IGCN_ERR_BAD_ARG,
                     "tc_gemm: bad destination segment %d", i);
        dst.p[i] = ds[i];
        dst.w[i] = (int)host_dst_widths[i];
        dst.ld[i] = (int)host_dst_strides[i];
        wsum += host_dst_widths[i];
    }
    IGCN_REQUIRE(wsum == N, IGCN_ERR_BAD_ARG, "tc_gemm: destination widths must add up to N");
    cudaStream_t st = (cudaStream_t)stream;
    if (N <= 64) return tc::launch_gemm<64>(a_hi, a_lo, lda, b_hi, b_lo, ldb, M, N, K, bias, (int)relu, dst, partials, S, st);
    return tc::launch_gemm<128>(a_hi, a_lo, lda, b_hi, b_lo, ldb, M, N, K, bias, (int)relu, dst, partials, S, st);
}

// host_jobs: njobs x 11 int64: {src, mask, hi, lo, rows, cols, ld_src, ld_dst, row_off, col_off, transpose}
extern "C" int igcn_tc_split(const int64_t* host_jobs, int64_t njobs, void* stream) {
    IGCN_REQUIRE(host_jobs && njobs >= 1 && njobs <= tc::MAX_JOBS, IGCN_ERR_BAD_ARG, "tc_split: 1..%d jobs", tc::MAX_JOBS);
    tc::SplitJobs jobs;
    int64_t max_tiles = 1;
    for (int i = 0; i < njobs; ++i) {
        const int64_t* h = host_jobs + 12 * i;
        tc::SplitJob& J = jobs.j[i];
        J.src = reinterpret_cast<const float*>(h[0]);
        J.mask = reinterpret_cast<const float*>(h[1]);
        J.hi = reinterpret_cast<float*>(h[2]);
        J.lo = reinterpret_cast<float*>(h[3]);
        J.rows = (int)h[4]; J.cols = (int)h[5]; J.ld_src = (int)h[6]; J.ld_dst = (int)h[7];
        J.row_off = (int)h[8]; J.col_off = (int)h[9]; J.transpose = (int)h[10];
        J.sub = reinterpret_cast<const float*>(h[11]);
        IGCN_REQUIRE(J.hi && J.lo && J.rows >= 0 && J.cols >= 0 && J.ld_dst > 0, IGCN_ERR_BAD_ARG, "tc_split: bad job %d", i);
        IGCN_REQUIRE(!J.mask || J.src, IGCN_ERR_BAD_ARG, "tc_split: job %d has a mask but no source", i);
        const int64_t t = (int64_t)((J.rows + 31) / 32) * ((J.cols + 31) / 32);
        if (t > max_tiles) max_tiles = t;
    }
    const int64_t cap = (int64_t)sm_count() * 16;
    dim3 grid((unsigned)(max_tiles < cap ? max_tiles : cap), (unsigned)njobs);
    igcn::launch_k(tc::tc_split_kernel, dim3(grid), dim3(256), 0, (cudaStream_t)stream, jobs);
    IGCN_CHECK_LAUNCH("tc_split");
    return IGCN_OK;
}

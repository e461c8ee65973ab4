// Warp-level tensor-core helpers shared by the SGCN encoder and the cross-attention kernels: mma.sync TF32 (m16n8k8 / m16n8k4),
// ldmatrix A-fragment loads of fp32 tiles, and the 2-instruction hi/lo split of the "3 x TF32" scheme
// (C += Al*Bh + Ah*Bl + Ah*Bh, fp32 accumulate; the dropped Al*Bl term is 2^-22 relative).
#pragma once
#include <stdint.h>

namespace igcn {
namespace mmau {

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint32_t tf32_rn(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return r;
}
// x = hi + lo EXACTLY: hi keeps the top 11 significand bits (a valid TF32 value), lo = x - hi is exact in fp32 (13 bits) and the
// tensor core reads its top 11 -- what is dropped is < 2^-21 |x|.  Two instructions (LOP3, FADD); cvt.rna.tf32 is emulated on
// sm_100a with a 4-instruction sequence per conversion (SASS: VIADD, FSETP, SEL, LOP3), which made the splits 12 % of the kernel.
__device__ __forceinline__ void split(float x, uint32_t& hi, uint32_t& lo) {
    hi = __float_as_uint(x) & 0xffffe000u;
    lo = __float_as_uint(x - __uint_as_float(hi));
}
__device__ __forceinline__ void mma_k8(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void mma_k4(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t b0) {
    asm volatile("mma.sync.aligned.m16n8k4.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a0), "r"(a1), "r"(b0));
}
// One 16 x 8 fp32 A tile (rows r..r+15, 8 consecutive floats) in the m16n8k8 fragment layout: an 8 x 4 fp32 block is an
// 8 x 8 b16 matrix to ldmatrix, and lane l receives element (row l/4, fp32 column l%4) -- exactly a0..a3.
__device__ __forceinline__ void ldmatrix_a(uint32_t (&a)[4], uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(a[0]), "=r"(a[1]), "=r"(a[2]), "=r"(a[3]) : "r"(addr));
}

}  // namespace mmau
}  // namespace igcn

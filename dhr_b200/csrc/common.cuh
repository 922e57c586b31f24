// common.cuh -- shared device/host helpers for the dhr_b200 kernels (sm_100a only).
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/dhr_b200.h"

namespace dhr {

// ---------------------------------------------------------------------------------------------
// host-side error plumbing
// ---------------------------------------------------------------------------------------------
void set_cuda_error(cudaError_t e, const char* what, const char* file, int line);

#define DHR_CUDA(expr)                                                        \
    do {                                                                      \
        cudaError_t _e = (expr);                                              \
        if (_e != cudaSuccess) {                                              \
            ::dhr::set_cuda_error(_e, #expr, __FILE__, __LINE__);             \
            return (_e == cudaErrorMemoryAllocation) ? DHR_ERR_NOMEM : DHR_ERR_CUDA; \
        }                                                                     \
    } while (0)

#define DHR_TRY(expr)                 \
    do {                              \
        int _s = (expr);              \
        if (_s != DHR_OK) return _s;  \
    } while (0)

static inline int64_t round_up(int64_t x, int64_t m) { return (x + m - 1) / m * m; }

// code values (slice-index codes stored in HBM).  A corpus slice whose G values are all zero
// stores CODE_EMPTY, a query slice that can never match stores CODE_NOMATCH: 0*x contributes
// nothing (densify_corpus.py:30-45 writes value 0 AND idx 0 for empty slices), so both are exact.
template <typename CodeT> struct CodeTraits;
template <> struct CodeTraits<uint8_t>  { static constexpr uint32_t kEmpty = 0xFFu,   kNoMatch = 0xFEu,   kMax = 0xFDu; };
template <> struct CodeTraits<uint16_t> { static constexpr uint32_t kEmpty = 0xFFFFu, kNoMatch = 0xFFFEu, kMax = 0xFFFDu; };

// ---------------------------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------------------------
// f32 += f16 * f16 in one instruction (PTX ISA 8.6 mixed-precision fma, SASS FHFMA on sm_100a):
// products of two fp16 numbers are exact in fp32, so this equals the reference's fp32 FMA on the
// fp32 copies of the same fp16 values (gip_retrieval.py:275,313).
__device__ __forceinline__ float fma_h_lo(uint32_t a, uint32_t b, float c) {
    float d;
    asm("{\n\t.reg .f16 al, ah, bl, bh;\n\tmov.b32 {al, ah}, %1;\n\tmov.b32 {bl, bh}, %2;\n\t"
        "fma.rn.f32.f16 %0, al, bl, %3;\n\t}"
        : "=f"(d) : "r"(a), "r"(b), "f"(c));
    return d;
}
__device__ __forceinline__ float fma_h_hi(uint32_t a, uint32_t b, float c) {
    float d;
    asm("{\n\t.reg .f16 al, ah, bl, bh;\n\tmov.b32 {al, ah}, %1;\n\tmov.b32 {bl, bh}, %2;\n\t"
        "fma.rn.f32.f16 %0, ah, bh, %3;\n\t}"
        : "=f"(d) : "r"(a), "r"(b), "f"(c));
    return d;
}
__device__ __forceinline__ float half_lo_to_float(uint32_t a) { return __half2float(__ushort_as_half((unsigned short)(a & 0xFFFFu))); }
__device__ __forceinline__ float half_hi_to_float(uint32_t a) { return __half2float(__ushort_as_half((unsigned short)(a >> 16))); }

// monotone map fp32 -> uint32 (larger float -> larger uint); -0.0 must be canonicalised by the caller
__device__ __forceinline__ uint32_t float_to_ordered(float f) {
    uint32_t b = __float_as_uint(f);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float ordered_to_float(uint32_t u) {
    return __uint_as_float((u & 0x80000000u) ? (u & 0x7FFFFFFFu) : ~u);
}
// 64-bit selection key: descending key order == (score desc, row asc)
__device__ __forceinline__ unsigned long long make_key(float score, uint32_t row) {
    return ((unsigned long long)float_to_ordered(score) << 32) | (unsigned long long)(0xFFFFFFFFu - row);
}
__device__ __forceinline__ float key_score(unsigned long long k) { return ordered_to_float((uint32_t)(k >> 32)); }
__device__ __forceinline__ uint32_t key_row(unsigned long long k) { return 0xFFFFFFFFu - (uint32_t)(k & 0xFFFFFFFFull); }

// ---- mbarrier / bulk-TMA PTX (cp.async.bulk -> SASS UBLKCP) ----------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
// Bounded wait: a protocol bug traps (launch fails with an error) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if (++spins > (1u << 26)) __trap();
    }
}
// 1-D bulk copy global -> shared, completion counted in bytes on `bar`
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(smem_dst)), "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

}  // namespace dhr

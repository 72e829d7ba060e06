// Shared device/host helpers for libpycgtool_b200 (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/pycgtool_b200.h"

#define CGB_CUDA_TRY(expr)                         \
    do {                                           \
        cudaError_t _e = (expr);                   \
        if (_e != cudaSuccess) return (int)_e;     \
    } while (0)

namespace cgb {

constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs

// ---------------------------------------------------------------- mbarrier + bulk-copy (TMA) PTX
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}

// Make barrier initialisation visible to the async (TMA) proxy.
__device__ __forceinline__ void fence_barrier_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}

// Announce `bytes` of asynchronous copies on the current phase without arriving.
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.expect_tx.relaxed.cta.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE;\n"
        "bra WAIT_LOOP;\n"
        "DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}

// 1-D bulk asynchronous copy global -> shared, completion signalled on an mbarrier (SASS: UBLKCP).
// dst and src must be 16-byte aligned and bytes a multiple of 16.
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(dst)),
        "l"(src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}

// Same with an L2 evict-first policy: the trajectory is streamed exactly once.
__device__ __forceinline__ void bulk_g2s_stream(void* dst, const void* src, uint32_t bytes, uint64_t* bar,
                                                uint64_t policy) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::
            "r"(smem_u32(dst)),
        "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
        : "memory");
}

__device__ __forceinline__ uint64_t policy_evict_first() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}

// Streaming (no L1 allocate) scalar store for write-once outputs.
__device__ __forceinline__ void st_stream(float* p, float v) { __stcs(p, v); }

// Packed float32x2 arithmetic (sm_100 FADD2): two frames per instruction, each half rounded to
// nearest exactly like the scalar operation.  ptxas contracts mul.rn.f32x2 + add.rn.f32x2 into one FFMA2
// (one rounding instead of two): mul2 below is the contraction-proof product, mulc2 the plain one.
__device__ __forceinline__ uint64_t pack2(float a, float b) {
    uint64_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
    return r;
}
__device__ __forceinline__ void unpack2(uint64_t v, float& a, float& b) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v));
}
__device__ __forceinline__ uint64_t sub2(uint64_t a, uint64_t b) {
    uint64_t r;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ uint64_t add2(uint64_t a, uint64_t b) {
    uint64_t r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) {
    uint64_t r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
// Packed product rounded to nearest, written as fma(a, b, +0) so that ptxas has no multiply to
// contract with the add that follows (it folds mul.rn.f32x2 -- and even fma(a, b, -0) -- plus
// add.rn.f32x2 into one FFMA2, which changes the rounding).  fma(a, b, +0) equals rn(a * b) except
// that a product of -0 becomes +0; the accumulator it is added to starts at +0 and can never be -0,
// so acc + (+0) == acc + (-0) and the sum is bit-identical to the reference's.
__device__ __forceinline__ uint64_t mul2(uint64_t a, uint64_t b) {
    uint64_t r;
    const uint64_t poszero2 = 0x0000000000000000ull;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(poszero2));
    return r;
}

// Plain packed product; ptxas may contract it with a following add2 into one FFMA2.
__device__ __forceinline__ uint64_t mulc2(uint64_t a, uint64_t b) {
    uint64_t r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ uint64_t splat2(float a) { return pack2(a, a); }

// ---------------------------------------------------------------- minimum image, one component
// Exactly the reference expression v -= L * rint(v / L) (mapping.py:457) with IEEE divide,
// round-half-even and un-fused multiply/subtract.  |v| <= 0.49 L  =>  rint(v / L) == 0 and the
// update is the identity, so the fast path is bit-identical.
__device__ __forceinline__ float mic1(float v, float L, float thr) {
    if (fabsf(v) > thr) {
        float k = rintf(__fdiv_rn(v, L));
        v = __fsub_rn(v, __fmul_rn(L, k));
    }
    return v;
}

// ---------------------------------------------------------------- staging spans with bulk copies
// Number of floats between a pointer and the 16-byte boundary below it.
__device__ __forceinline__ uint32_t head_floats(const float* p) {
    return (uint32_t)((reinterpret_cast<uintptr_t>(p) & 15u) >> 2);
}

// Stage `nfl` floats starting at `src` into `slot` so that float i lands at slot[head + i].
// Interior runs go through one bulk copy from the 16-byte-aligned address below `src`; a run that
// would touch bytes outside [lo, hi) (only the first / last bytes of the whole array) is copied by
// the producer warp with plain loads.  Returns the number of bytes the mbarrier must expect.
__device__ __forceinline__ uint32_t span_bytes(const float* src, uint32_t nfl, uintptr_t lo, uintptr_t hi,
                                               bool* edge) {
    const uintptr_t a = reinterpret_cast<uintptr_t>(src);
    const uintptr_t a0 = a & ~uintptr_t(15);
    const uintptr_t e1 = (a + 4u * (uintptr_t)nfl + 15u) & ~uintptr_t(15);
    *edge = (a0 < lo) || (e1 > hi);
    return *edge ? 0u : (uint32_t)(e1 - a0);
}


}  // namespace cgb

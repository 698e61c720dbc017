// Shared device helpers for the B200 (sm_100a) acoustic front end.
//  * packed fp32x2 arithmetic (FADD2/FMUL2/FFMA2): one issue slot per complex add and two per
//    complex multiply; ptxas folds the half swaps / sign flips / scalar broadcasts written
//    below into operand modifiers (R.F32x2.LO_HI.NP, R.F32), so they cost no instructions.
//  * mbarrier + 1-D bulk async copy (TMA, cp.async.bulk -> SASS UBLKCP) for waveform tiles.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace b200fe {

typedef unsigned long long u64;

__device__ __forceinline__ u64 pk(float2 a) { u64 d; asm("mov.b64 %0, {%1,%2};" : "=l"(d) : "f"(a.x), "f"(a.y)); return d; }
__device__ __forceinline__ float2 upk(u64 v) { float2 r; asm("mov.b64 {%0,%1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v)); return r; }
__device__ __forceinline__ float2 add2(float2 a, float2 b) { u64 d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(pk(a)), "l"(pk(b))); return upk(d); }
__device__ __forceinline__ float2 sub2(float2 a, float2 b) { u64 d; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(pk(a)), "l"(pk(b))); return upk(d); }
__device__ __forceinline__ float2 mul2(float2 a, float2 b) { u64 d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(pk(a)), "l"(pk(b))); return upk(d); }
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) { u64 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(pk(a)), "l"(pk(b)), "l"(pk(c))); return upk(d); }
__device__ __forceinline__ float2 bc(float s) { return make_float2(s, s); }
__device__ __forceinline__ float2 swp(float2 a) { return make_float2(a.y, a.x); }

// complex helpers on (re, im) pairs
__device__ __forceinline__ float2 c_amjb(float2 a, float2 b) { return fma2(swp(b), make_float2(1.f, -1.f), a); }  // a - j*b
__device__ __forceinline__ float2 c_apjb(float2 a, float2 b) { return fma2(swp(b), make_float2(-1.f, 1.f), a); }  // a + j*b
__device__ __forceinline__ float2 c_mul(float2 z, float c, float d) {                                             // z * (c + j d)
    float2 t = mul2(z, bc(c));
    return fma2(swp(z), make_float2(-d, d), t);
}

// Radix-4 forward butterfly (omega_4 = -j), in place: (x0,x1,x2,x3) -> (X0,X1,X2,X3).  8 packed instrs.
__device__ __forceinline__ void bfly4(float2& x0, float2& x1, float2& x2, float2& x3)
{
    float2 s0 = add2(x0, x2), d0 = sub2(x0, x2), s1 = add2(x1, x3), d1 = sub2(x1, x3);
    x0 = add2(s0, s1);
    x2 = sub2(s0, s1);
    x1 = c_amjb(d0, d1);
    x3 = c_apjb(d0, d1);
}

// 16-point forward DFT on registers, natural order in -> natural order out (4x4 Cooley-Tukey,
// all indices static).  Inputs that are compile-time zero are pruned by the compiler.
__device__ __forceinline__ void dft16(float2 (&v)[16])
{
    const float C1 = 0.92387953251128673848f, S1 = 0.38268343236508978178f, R2 = 0.70710678118654752440f;
    // step 1: radix-4 over n_b for each n_a (elements n_a, n_a+4, n_a+8, n_a+12) -> B[n_a][kl] lives in v[n_a+4*kl]
#pragma unroll
    for (int na = 0; na < 4; ++na) bfly4(v[na], v[na + 4], v[na + 8], v[na + 12]);
    // step 2: twiddle B[n_a][kl] *= w16^(n_a*kl)   (w16 = exp(-2 pi j/16))
    v[1 + 4] = c_mul(v[1 + 4], C1, -S1);    // w^1
    v[1 + 8] = c_mul(v[1 + 8], R2, -R2);    // w^2
    v[1 + 12] = c_mul(v[1 + 12], S1, -C1);  // w^3
    v[2 + 4] = c_mul(v[2 + 4], R2, -R2);    // w^2
    v[2 + 8] = make_float2(v[2 + 8].y, -v[2 + 8].x);  // w^4 = -j
    v[2 + 12] = c_mul(v[2 + 12], -R2, -R2); // w^6
    v[3 + 4] = c_mul(v[3 + 4], S1, -C1);    // w^3
    v[3 + 8] = c_mul(v[3 + 8], -R2, -R2);   // w^6
    v[3 + 12] = c_mul(v[3 + 12], -C1, S1);  // w^9
    // step 3: radix-4 over n_a for each kl: inputs v[0+4kl..3+4kl] -> A[kl + 4*kh] ends in v[kh + 4*kl]
#pragma unroll
    for (int kl = 0; kl < 4; ++kl) bfly4(v[4 * kl], v[4 * kl + 1], v[4 * kl + 2], v[4 * kl + 3]);
    // un-transpose: A[kl + 4 kh] sits in v[kh + 4 kl]  ->  want v[k] = A[k]
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = a + 1; b < 4; ++b) { float2 t = v[a + 4 * b]; v[a + 4 * b] = v[b + 4 * a]; v[b + 4 * a] = t; }
}

// ---- mbarrier / bulk copy ------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}"
        ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// 1-D TMA bulk copy global -> shared, completion signalled on an mbarrier (bytes % 16 == 0,
// src/dst 16-byte aligned).
__device__ __forceinline__ void tma_load_1d(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// Programmatic dependent launch (sm_90+): the builder -> fused -> post-pass chain of a call is launched with programmatic stream
// serialisation, so a kernel's CTAs may become resident, and run the part of their prologue that touches only plan constants, while
// the preceding kernel drains.  pdl_wait() returns when the preceding kernel has completed and its writes are visible: every read of
// caller data and every write comes after it.  Both are no-ops for a normally launched kernel.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

__device__ __forceinline__ float fast_log(float x)
{
    // lg2.approx: max rel err 2^-22 on the mantissa path; |abs err| <= ~4e-7 near 1, <= 3 ulp elsewhere.
    float r; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r * 0.69314718055994530942f;
}

}  // namespace b200fe

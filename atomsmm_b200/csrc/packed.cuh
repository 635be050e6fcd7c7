// Two-wide fp32 value type over the packed f32x2 instructions of sm_100a (FADD2 / FMUL2 / FFMA2: one
// instruction issue, two results).  The pair tiles and the neighbour-list sweep are bound by instruction
// issue, so they do their distance and force arithmetic on F2.  The *_rn intrinsics are never contracted
// by the compiler: fused multiply-adds are written out as fma2().  On the host (tests/native/) the type
// falls back to component-wise fmaf, so the packed restatements can be checked without a GPU.
#pragma once

#include <cuda_runtime.h>
#include <math.h>

#ifdef __CUDACC__
#define B2_HD __host__ __device__ __forceinline__
#else
#define B2_HD inline
#endif

struct F2 {
    float2 v;
};
B2_HD F2 f2(float a) { return F2{make_float2(a, a)}; }
B2_HD F2 f2(float a, float b) { return F2{make_float2(a, b)}; }
B2_HD F2 operator*(F2 a, F2 b) {
#ifdef __CUDA_ARCH__
    return F2{__fmul2_rn(a.v, b.v)};
#else
    return F2{make_float2(a.v.x*b.v.x, a.v.y*b.v.y)};
#endif
}
B2_HD F2 operator+(F2 a, F2 b) {
#ifdef __CUDA_ARCH__
    return F2{__fadd2_rn(a.v, b.v)};
#else
    return F2{make_float2(a.v.x + b.v.x, a.v.y + b.v.y)};
#endif
}
B2_HD F2 fma2(F2 a, F2 b, F2 c) {
#ifdef __CUDA_ARCH__
    return F2{__ffma2_rn(a.v, b.v, c.v)};
#else
    return F2{make_float2(fmaf(a.v.x, b.v.x, c.v.x), fmaf(a.v.y, b.v.y, c.v.y))};
#endif
}
B2_HD F2 max0(F2 a) { return F2{make_float2(fmaxf(a.v.x, 0.f), fmaxf(a.v.y, 0.f))}; }
B2_HD float b2_rsqrt_approx(float x) {
#ifdef __CUDA_ARCH__
    float y;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
#else
    return 1.0f/sqrtf(x);
#endif
}


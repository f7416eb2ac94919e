// packed.cuh -- packed fp32x2 arithmetic helpers (FFMA2 / FADD2 / FMUL2 on sm_100a) shared by the engines
#pragma once
#include <cuda_runtime.h>

namespace b2fwi {

// ---- packed fp32x2 helpers (FFMA2 / FADD2 / FMUL2 on sm_100a)
static __device__ __forceinline__ float4 z4() { return make_float4(0.f, 0.f, 0.f, 0.f); }
static __device__ __forceinline__ float2 lo2(float4 a) { return make_float2(a.x, a.y); }
static __device__ __forceinline__ float2 hi2(float4 a) { return make_float2(a.z, a.w); }
static __device__ __forceinline__ float4 mk4(float2 l, float2 h) { return make_float4(l.x, l.y, h.x, h.y); }
static __device__ __forceinline__ float4 add4(float4 a, float4 b)
{
    return mk4(__fadd2_rn(lo2(a), lo2(b)), __fadd2_rn(hi2(a), hi2(b)));
}
static __device__ __forceinline__ float4 fma4s(float s, float4 a, float4 c)   // s*a + c
{
    const float2 ss = make_float2(s, s);
    return mk4(__ffma2_rn(ss, lo2(a), lo2(c)), __ffma2_rn(ss, hi2(a), hi2(c)));
}
static __device__ __forceinline__ float4 fma4(float4 a, float4 b, float4 c)   // a*b + c
{
    return mk4(__ffma2_rn(lo2(a), lo2(b), lo2(c)), __ffma2_rn(hi2(a), hi2(b), hi2(c)));
}
static __device__ __forceinline__ float4 mul4(float4 a, float4 b)
{
    return mk4(__fmul2_rn(lo2(a), lo2(b)), __fmul2_rn(hi2(a), hi2(b)));
}
static __device__ __forceinline__ float4 mul4s(float s, float4 a)
{
    const float2 ss = make_float2(s, s);
    return mk4(__fmul2_rn(ss, lo2(a)), __fmul2_rn(ss, hi2(a)));
}
static __device__ __forceinline__ float rcp_approx(float x)
{
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}


}  // namespace b2fwi

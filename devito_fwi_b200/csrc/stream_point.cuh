// stream_point.cuh -- the per-point arithmetic shared by every streaming kernel variant (register-staged,
// cp.async, TMA): identical operation order => bitwise identical sweeps whichever variant runs.
#pragma once
#include "common.cuh"
#include "packed.cuh"
#include "stream_kernels.cuh"

namespace b2fwi {

static __device__ __forceinline__ float4 ld4(const float *p) { return *reinterpret_cast<const float4 *>(p); }
static __device__ __forceinline__ void st4(float *p, float4 v) { *reinterpret_cast<float4 *>(p) = v; }
static __device__ __forceinline__ float4 zero4() { return make_float4(0.f, 0.f, 0.f, 0.f); }
// u.dt2 from three time levels; one fixed operation order everywhere so that a checkpointed
// gradient (u.dt2 stored by the recompute sweep) is bitwise identical to the full-history one.
static __device__ __forceinline__ float d2u_of(float um, float uc, float up, float inv_dt2)
{
    return __fmul_rn(__fadd_rn(__fmaf_rn(-2.f, uc, um), up), inv_dt2);
}

// Sparse operators, one definition for the stand-alone kernels and the service warps of the TMA sweeps (same
// operation order => the fused and the three-launch step are bitwise identical).
// field[cell] += sum_j w_j * vals[pt_j] * dt^2 * vp[cell]^2, contributions in ascending point order
// (operators.py:134,221: expr = src * s**2 / m evaluated at each corner's own vp).
static __device__ __forceinline__ float inject_term(float w, float val, float dt, float v)
{
    return __fmul_rn(__fmul_rn(__fmul_rn(__fmul_rn(__fmul_rn(w, val), dt), dt), v), v);
}
static __device__ __forceinline__ float inject_cell(float f, int j0, int j1, const float *__restrict__ contrib_w,
                                                   const int32_t *__restrict__ contrib_pt,
                                                   const float *__restrict__ vals, float dt, float v)
{
    for (int j = j0; j < j1; j++) f = __fadd_rn(f, inject_term(contrib_w[j], vals[contrib_pt[j]], dt, v));
    return f;
}
// out[p] = sum_c w_c * field[c]  (operators.py:137,176)
static __device__ __forceinline__ float interp_point(const float *__restrict__ field, int64_t p, int ncorner,
                                                    const int64_t *__restrict__ corner_off,
                                                    const float *__restrict__ corner_w)
{
    float sum = 0.f;
    for (int c = 0; c < ncorner; c++) {
        const int64_t off = corner_off[p * ncorner + c];
        if (off >= 0) sum += corner_w[p * ncorner + c] * field[off];
    }
    return sum;
}
// imaging by parts: the sweep formed v.dt2 before the injection, the injected increment is added afterwards
static __device__ __forceinline__ float img_inject_fix(float g, float hist, float f_new, float f_old, float inv_dt2)
{
    return __fmaf_rn(-hist, __fmul_rn(__fsub_rn(f_new, f_old), inv_dt2), g);
}

// Staged-tile accessors: a generic pointer (register-staged / cp.async kernels) or a 32-bit shared-window address.
struct SAddr { uint32_t a; };
static __device__ __forceinline__ float4 ldt(const float *p, int off) { return ld4(p + off); }
static __device__ __forceinline__ float4 ldt(SAddr s, int off)
{
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(s.a + 4u * (uint32_t)off));
    return v;
}

// The point update shared by all streaming kernels (identical arithmetic => identical results whichever
// kernel variant computes a sweep): packed fp32x2, three independent accumulation chains.
//   q[0..NQ-1]: register pipeline along the plane axis (centre at QC); ctr: this thread's float4 in the staged tile.
//   The pipeline is addressed circularly: plane offset i in -R..R lives in q[(j + R + i) % NQ] (j: rotation).
//   W: anything carrying the Laplacian weights (StepArgs in constant memory, or StencilW pinned in registers).
//   MASKZ = false: the caller guarantees c2 == 0 and u == 0 beyond nz (TMA zero fill), which makes the update 0 there.
// (split in two so that a kernel short of registers can fetch prev / c1 / c2 AFTER the Laplacian: point_laplacian +
// point_finish is point_update, operation for operation)
// Free surface (operators.py:8-35): in the top rows a reference u[z - k] becomes sign(z - k) * u[|z - k|] - the
// antisymmetric mirror about index 0, whose own value counts as zero when it is reached through a negative offset.
// zl[t] holds u at z = fsz - ZH + t (fsz: global z of the thread's quad, 0 / 4 / 8): positions <= 0 are rewritten.
template <int F, int ZH>
static __device__ __forceinline__ void mirror_top(float *zl)
{
    if (F <= ZH) {
#pragma unroll
        for (int t = 0; t < ZH - F; t++) zl[t] = -zl[2 * ZH - 2 * F - t];
        zl[ZH - F] = 0.f;
    }
}

// fsz: global z index of the quad when the model has a free surface and the quad is within reach of it (<= R), else -1
template <int R, int NDIM, int SW, class W = StepArgs, class TP = const float *>
static __device__ __forceinline__ float4 point_laplacian(const W &a, const float4 *q, int j, TP ctr, int fsz = -1)
{
    constexpr int RZ4 = (R + 3) / 4, ZH = 4 * RZ4;
    constexpr int NQ = (NDIM == 3) ? 2 * R + 1 : 1;
    constexpr int QC = (NDIM == 3) ? R : 0;
    const float4 C = q[(j + QC) % NQ];
    float4 lp = fma4s(a.c0, C, mul4s(a.c0_lo, C));          // centre weight as exact hi + lo (see api.cu)
    if (NDIM == 3) {
#pragma unroll
        for (int k = 1; k <= R; k++) lp = fma4s(a.cp[k], add4(q[(j + QC + k) % NQ], q[(j + QC - k + NQ) % NQ]), lp);
    }
    float4 lr = mul4s(a.cr[1], add4(ldt(ctr, SW), ldt(ctr, -SW)));
#pragma unroll
    for (int k = 2; k <= R; k++) lr = fma4s(a.cr[k], add4(ldt(ctr, k * SW), ldt(ctr, -k * SW)), lr);
    float zl[ZH + 4 + ZH];
#pragma unroll
    for (int i = 0; i < RZ4; i++) {
        const float4 Lq = ldt(ctr, -ZH + 4 * i), Rq = ldt(ctr, 4 + 4 * i);
        zl[4 * i + 0] = Lq.x; zl[4 * i + 1] = Lq.y; zl[4 * i + 2] = Lq.z; zl[4 * i + 3] = Lq.w;
        zl[ZH + 4 + 4 * i + 0] = Rq.x; zl[ZH + 4 + 4 * i + 1] = Rq.y;
        zl[ZH + 4 + 4 * i + 2] = Rq.z; zl[ZH + 4 + 4 * i + 3] = Rq.w;
    }
    zl[ZH + 0] = C.x; zl[ZH + 1] = C.y; zl[ZH + 2] = C.z; zl[ZH + 3] = C.w;
    if (fsz >= 0) {          // a few threads per row, and only in models with a free surface
        if (fsz == 0) mirror_top<0, ZH>(zl);
        else if (fsz == 4) mirror_top<4, ZH>(zl);
        else mirror_top<8, ZH>(zl);
    }
    // z direction in packed pairs. Even offsets k pair up naturally: (o0,o1) = (z[k],z[k+1]) + (z[-k],z[1-k]) are
    // aligned register pairs. For odd k those pairs straddle registers (two MOVs each), so the sums are formed
    // for the aligned output pairs (o-1,o0), (o1,o2), (o3,o4) instead - three packed adds, outer lanes unused -
    // and folded into (o0..o3) by four scalar adds at the end.
#define ZP(i) make_float2(zl[ZH + (i)], zl[ZH + (i) + 1])
    float2 l01 = make_float2(0.f, 0.f), l23 = l01, oa = l01, ob = l01, oc = l01;
#pragma unroll
    for (int k = 1; k <= R; k++) {
        const float2 ck = make_float2(a.cz[k], a.cz[k]);
        if (k & 1) {
            const float2 sa = __fadd2_rn(ZP(k - 1), ZP(-k - 1)), sb = __fadd2_rn(ZP(k + 1), ZP(1 - k)),
                         sc = __fadd2_rn(ZP(k + 3), ZP(3 - k));
            if (k == 1) { oa = __fmul2_rn(ck, sa); ob = __fmul2_rn(ck, sb); oc = __fmul2_rn(ck, sc); }
            else { oa = __ffma2_rn(ck, sa, oa); ob = __ffma2_rn(ck, sb, ob); oc = __ffma2_rn(ck, sc, oc); }
        } else {
            const float2 s0 = __fadd2_rn(ZP(k), ZP(-k)), s2 = __fadd2_rn(ZP(k + 2), ZP(2 - k));
            if (k == 2) { l01 = __fmul2_rn(ck, s0); l23 = __fmul2_rn(ck, s2); }
            else { l01 = __ffma2_rn(ck, s0, l01); l23 = __ffma2_rn(ck, s2, l23); }
        }
    }
#undef ZP
    if (R >= 2) {
        l01.x = __fadd_rn(l01.x, oa.y); l01.y = __fadd_rn(l01.y, ob.x);
        l23.x = __fadd_rn(l23.x, ob.y); l23.y = __fadd_rn(l23.y, oc.x);
    } else {
        l01 = make_float2(oa.y, ob.x); l23 = make_float2(ob.y, oc.x);
    }
    return add4(add4(lp, lr), mk4(l01, l23));
}

template <bool MASKZ = true>
static __device__ __forceinline__ float4 point_finish(float4 C, float4 lap, float4 prev, float4 c1, float4 c2, int zvalid)
{
    // u+ = u + c1 (u - u-) + c2 L(u)
    const float4 t = fma4(c1, add4(C, make_float4(-prev.x, -prev.y, -prev.z, -prev.w)), C);
    float4 o = fma4(c2, lap, t);
    if (MASKZ && zvalid < 4) {            // only the last, partial quad of a row: keep the pitch padding zero
        if (zvalid < 2) o.y = 0.f;
        if (zvalid < 3) o.z = 0.f;
        o.w = 0.f;
    }
    return o;
}

template <int R, int NDIM, int SW, bool MASKZ = true, class W = StepArgs, class TP = const float *>
static __device__ __forceinline__ float4 point_update(const W &a, const float4 *q, int j, TP ctr,
                                                     float4 prev, float4 c1, float4 c2, int zvalid, int fsz = -1)
{
    constexpr int NQ = (NDIM == 3) ? 2 * R + 1 : 1;
    constexpr int QC = (NDIM == 3) ? R : 0;
    const float4 lap = point_laplacian<R, NDIM, SW, W, TP>(a, q, j, ctr, fsz);
    return point_finish<MASKZ>(q[(j + QC) % NQ], lap, prev, c1, c2, zvalid);
}

// Laplacian weights held in registers (filled by the kernel from a source ptxas cannot rematerialise)
template <int R>
struct StencilW {
    float c0, c0_lo, cp[R + 1], cr[R + 1], cz[R + 1];
};

// packed forms of d2u_of / the imaging update (FFMA2 etc. round exactly like their scalar counterparts)
static __device__ __forceinline__ float4 d2u4(float4 um, float4 uc, float4 up, float inv_dt2)
{
    return mul4s(inv_dt2, add4(fma4s(-2.f, uc, um), up));
}
static __device__ __forceinline__ float4 img4(float4 g, float4 d2, float4 v)     // grad += -u.dt2 * v
{
    return fma4(make_float4(-d2.x, -d2.y, -d2.z, -d2.w), v, g);
}

}  // namespace b2fwi

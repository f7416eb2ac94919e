// stream_kernels.cu -- "streaming" engine: one launch per time step, wavefields in HBM.
//
// Used for every 3-D model and for 2-D models too large for the SM-resident engine.
// Replaces the generated C of Devito's Forward / Gradient / Adjoint operators
// (reference: seismic/acoustic/operators.py:59-95 iso_stencil, :98-140, :143-180, :183-225).
//
// step_kernel<R, NDIM, IMG>: 2.5-D sweep. A CTA owns a (16 rows x 64 z) tile and streams along
// the plane (slow, x) axis of a 3-D grid with a register pipeline of 2R+1 float4 per thread
// (128-bit coalesced HBM loads along the contiguous z axis); the current plane's tile plus its
// row/z halos is staged in double-buffered shared memory (one __syncthreads per plane); next
// plane's halo and pipeline head are prefetched into registers while the current plane computes.
// The same kernel fuses, per point: the OT2 update, the zero-lag imaging condition
// (grad -= u.dt2 * v), the source-illumination accumulation and the u.dt2 history store.
#include <stdlib.h>

#include "common.cuh"
#include "packed.cuh"
#include "stream_kernels.cuh"
#include "stream_point.cuh"

namespace b2fwi {

// TZQ x TR: tile shape in float4 columns x rows (threads = TZQ*TR); MINB: minimum resident CTAs per SM.
template <int R, int NDIM, int IMG, int MINB, int TZQ = 16, int TR = 16>
__global__ void __launch_bounds__(TZQ *TR, MINB) step_kernel(const __grid_constant__ StepArgs a)
{
    constexpr int NT = TZQ * TR;           // threads per CTA
    constexpr int TZ = 4 * TZQ;            // z cells per tile
    constexpr int RZ4 = (R + 3) / 4;       // z-halo width in float4
    constexpr int ZH = 4 * RZ4;            // z-halo width in floats
    constexpr int SW = TZ + 2 * ZH;        // shared row width (floats)
    constexpr int SROWS = TR + 2 * R;
    constexpr int NHALO = 2 * R * TZQ + TR * 2 * RZ4;
    constexpr int NH = (NHALO + NT - 1) / NT;
    constexpr int NQ = (NDIM == 3) ? 2 * R + 1 : 1;
    constexpr int QC = (NDIM == 3) ? R : 0;

    __shared__ __align__(16) float tile[2][SROWS][SW];

    const int tz = threadIdx.x, tr = threadIdx.y, tid = tr * TZQ + tz;
    const int ztile0 = blockIdx.x * TZ;
    const int z0 = ztile0 + tz * 4;
    const int r0 = blockIdx.y * TR;
    const int r = r0 + tr;
    const bool active = (r < a.nr) && (z0 < a.nz);
    const int zvalid = min(4, a.nz - z0);      // valid lanes of this thread's float4
    const int p_begin = (NDIM == 3) ? (int)blockIdx.z * a.chunk : 0;
    const int p_end = (NDIM == 3) ? min(p_begin + a.chunk, a.np) : 1;
    const int64_t H = a.halo;
    // offset of this thread's float4 in plane 0
    const int64_t own0 = (NDIM == 3 ? H * a.sp : 0) + (int64_t)(r + H) * a.sr + (z0 + H);
    // every global operand is addressed as (uniform base) + 32-bit float4 index: a slice has < 2^32 float4
    // (one IMAD.WIDE per access instead of 64-bit add chains; the kernel is issue-sensitive)
    const uint32_t own4 = (uint32_t)(own0 >> 2);
    const uint32_t sp4 = (uint32_t)(a.sp >> 2);
#define F4(ptr) reinterpret_cast<const float4 *>(ptr)
#define F4W(ptr) reinterpret_cast<float4 *>(ptr)

    // ---- halo work items of this thread (fixed across planes)
    uint32_t hoff[NH];     // float4 index inside plane 0, or 0xffffffff
    int hsm[NH];           // shared offset (floats) inside one buffer, or -1
#pragma unroll
    for (int i = 0; i < NH; i++) {
        const int h = tid + i * NT;
        int srow = -1, scol = 0, gz = 0;
        if (h < 2 * R * TZQ) {
            const int hr = h / TZQ, hz = h % TZQ;
            srow = (hr < R) ? hr : hr + TR;
            scol = ZH + 4 * hz;
            gz = ztile0 + 4 * hz;
        } else if (h < NHALO) {
            const int j = h - 2 * R * TZQ;
            const int row = j / (2 * RZ4), c = j % (2 * RZ4);
            srow = R + row;
            if (c < RZ4) { scol = 4 * c; gz = ztile0 - ZH + 4 * c; }
            else { scol = ZH + TZ + 4 * (c - RZ4); gz = ztile0 + TZ + 4 * (c - RZ4); }
        }
        hsm[i] = (srow >= 0) ? srow * SW + scol : -1;
        const int gr = r0 - R + srow;
        const bool ok = (srow >= 0) && gr >= 0 && gr < a.nr && gz >= 0 && gz < a.nz;
        hoff[i] = ok ? (uint32_t)(((NDIM == 3 ? H * a.sp : 0) + (int64_t)(gr + H) * a.sr + (gz + H)) >> 2) : 0xffffffffu;
    }

    // undamped interior (coef[0] == 1 exactly): this thread never reads c1 for planes inside the box
    const int blo_p = __ldg(a.box + 0), bhi_p = __ldg(a.box + 1);
    const bool in_box = __ldg(a.box + 6) == 1 && r >= __ldg(a.box + 2) && r < __ldg(a.box + 3) &&
                        z0 >= __ldg(a.box + 4) && z0 + 3 < __ldg(a.box + 5);
    const float4 one4 = make_float4(1.f, 1.f, 1.f, 1.f);

    // ---- prologue: register pipeline along the plane axis, halo of the first plane
    float4 q[NQ];
    if (NDIM == 3) {
#pragma unroll
        for (int i = 0; i < NQ; i++) {
            const int p = p_begin - R + i;
            q[i] = (active && p >= 0 && p < a.np) ? F4(a.cur)[own4 + (uint32_t)p * sp4] : zero4();
        }
    } else {
        q[0] = active ? F4(a.cur)[own4] : zero4();
    }
    float4 hreg[NH];
#pragma unroll
    for (int i = 0; i < NH; i++)
        hreg[i] = (hoff[i] != 0xffffffffu) ? F4(a.cur)[hoff[i] + (NDIM == 3 ? (uint32_t)p_begin * sp4 : 0u)] : zero4();

    // plane loop; for R <= 4 it is unrolled NQ times so that the register pipeline rotates through
    // compile-time slots instead of being shifted (2R float4 moves per plane otherwise)
    constexpr int UNR = (NDIM == 3 && R <= 4) ? NQ : 1;
    uint32_t pofs_run = (NDIM == 3) ? (uint32_t)p_begin * sp4 : 0u;    // plane offset (float4), advanced by one plane per iteration
    for (int pb = p_begin; pb < p_end; pb += UNR) {
#pragma unroll
        for (int j = 0; j < UNR; j++) {
            const int p = pb + j;
            if (p < p_end) {
                const int buf = (p - p_begin) & 1;
                float *tb = &tile[buf][0][0];
                const uint32_t pofs = pofs_run;
                const uint32_t idx = own4 + pofs;
                pofs_run += sp4;
                // stage plane p
                st4(tb + (R + tr) * SW + ZH + 4 * tz, q[(j + QC) % NQ]);
#pragma unroll
                for (int i = 0; i < NH; i++)
                    if (hsm[i] >= 0) st4(tb + hsm[i], hreg[i]);

                // pointwise operands of plane p
                float4 prev, c1, c2, g4, h0, h1, h2, il;      // only defined (and only used) by active threads
                if (active) {
                    prev = F4(a.prev)[idx];
                    c1 = (in_box && p >= blo_p && p < bhi_p) ? one4 : __ldg(F4(a.c1) + idx);
                    c2 = __ldg(F4(a.c2) + idx);
                    if (IMG != 0) {
                        g4 = F4(a.grad)[idx];
                        h1 = __ldg(F4(a.h1) + idx);
                        if (IMG == 1) {
                            h0 = __ldg(F4(a.h0) + idx);
                            h2 = __ldg(F4(a.h2) + idx);
                        }
                    }
                    if (a.illum) il = F4(a.illum)[idx];
                }
                // prefetch for plane p+1
                float4 qn = zero4();
                if (p + 1 < p_end) {
                    if (NDIM == 3) {
                        const int pn = p + R + 1;
                        if (active && pn < a.np) qn = F4(a.cur)[idx + (uint32_t)(R + 1) * sp4];
                    }
#pragma unroll
                    for (int i = 0; i < NH; i++)
                        hreg[i] = (hoff[i] != 0xffffffffu) ? F4(a.cur)[hoff[i] + pofs + sp4] : zero4();
                }
                __syncthreads();

                if (active) {
                    const float4 C = q[(j + QC) % NQ];
                    const float4 o = point_update<R, NDIM, SW>(a, q, j, tb + (R + tr) * SW + ZH + 4 * tz, prev, c1, c2, zvalid,
                                                               (a.fs && z0 <= R) ? z0 : -1);
                    F4W(a.out)[idx] = o;
                    if (IMG == 1) F4W(a.grad)[idx] = img4(g4, d2u4(h0, h1, h2, a.inv_dt2), C);
                    if (IMG == 2) F4W(a.grad)[idx] = img4(g4, h1, a.hist_uv ? d2u4(prev, C, o, a.inv_dt2) : C);
                    if (a.illum) F4W(a.illum)[idx] = fma4(C, C, il);
                    if (a.d2u) {
                        float4 d = d2u4(prev, C, o, a.inv_dt2);
                        if (zvalid < 4) {
                            if (zvalid < 2) d.y = 0.f;
                            if (zvalid < 3) d.z = 0.f;
                            d.w = 0.f;
                        }
                        F4W(a.d2u)[idx] = d;
                    }
                }
                if (NDIM == 3) {
                    if (UNR == NQ) {
                        q[j % NQ] = qn;                     // the oldest plane's slot receives plane p+R+1
                    } else {
#pragma unroll
                        for (int i = 0; i < NQ - 1; i++) q[i] = q[i + 1];
                        q[NQ - 1] = qn;
                    }
                }
            }
        }
    }
}
#undef F4
#undef F4W

// ------------------------------------------------------------------------------------------------
// 3-D variant with an asynchronous-copy pipeline (cp.async, three stages): the halo of plane p+2 goes
// straight into the shared tile ring and the pointwise operands (u[t-1], the two coefficients and, for the
// imaging sweep, grad and u.dt2) of plane p+2 into thread-private shared slots while plane p computes, so no
// HBM latency sits between the per-plane barrier and the arithmetic. IMG: 0 forward, 2 imaging from u.dt2.
static __device__ __forceinline__ void cp_async16(uint32_t dst, const void *src, bool valid)
{
    const int sz = valid ? 16 : 0;        // src-size 0: 16 bytes of zeros are written, nothing is read
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
}
static __device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
static __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <int R, int IMG>
__global__ void __launch_bounds__(256, 2) step3d_async_kernel(const __grid_constant__ StepArgs a)
{
    constexpr int TR = 16, RZ4 = (R + 3) / 4, ZH = 4 * RZ4, SW = 64 + 2 * ZH, SROWS = TR + 2 * R;
    constexpr int NHALO = 2 * R * 16 + TR * 2 * RZ4, NH = (NHALO + 255) / 256;
    constexpr int NQ = 2 * R + 1, QC = R;
    constexpr int NB = 3;                               // pipeline stages
    constexpr int NAUX = (IMG == 2) ? 5 : 3;            // prev, c1, c2 [, grad, u.dt2]

    extern __shared__ __align__(16) float dsm[];
    float *tiles = dsm;                                               // [NB][SROWS][SW]
    float4 *aux = reinterpret_cast<float4 *>(dsm + NB * SROWS * SW);  // [NB][NAUX][256], thread-private slots

    const int tz = threadIdx.x, tr = threadIdx.y, tid = tr * 16 + tz;
    const int ztile0 = blockIdx.x * 64;
    const int z0 = ztile0 + tz * 4;
    const int r0 = blockIdx.y * TR;
    const int r = r0 + tr;
    const bool active = (r < a.nr) && (z0 < a.nz);
    const int zvalid = min(4, a.nz - z0);      // valid lanes of this thread's float4
    const int p_begin = (int)blockIdx.z * a.chunk;
    const int p_end = min(p_begin + a.chunk, a.np);
    const int64_t H = a.halo;
    const int64_t own0 = H * a.sp + (int64_t)(r + H) * a.sr + (z0 + H);
    const uint32_t tiles_s = (uint32_t)__cvta_generic_to_shared(tiles);
    const uint32_t aux_s = (uint32_t)__cvta_generic_to_shared(aux);

    int64_t hoff[NH];
    int hsm[NH];
#pragma unroll
    for (int i = 0; i < NH; i++) {
        const int h = tid + i * 256;
        int srow = -1, scol = 0, gz = 0;
        if (h < 2 * R * 16) {
            const int hr = h >> 4, hz = h & 15;
            srow = (hr < R) ? hr : hr + TR;
            scol = ZH + 4 * hz;
            gz = ztile0 + 4 * hz;
        } else if (h < NHALO) {
            const int j = h - 2 * R * 16;
            const int row = j / (2 * RZ4), c = j % (2 * RZ4);
            srow = R + row;
            if (c < RZ4) { scol = 4 * c; gz = ztile0 - ZH + 4 * c; }
            else { scol = ZH + 64 + 4 * (c - RZ4); gz = ztile0 + 64 + 4 * (c - RZ4); }
        }
        hsm[i] = (srow >= 0) ? srow * SW + scol : -1;
        const int gr = r0 - R + srow;
        const bool ok = (srow >= 0) && gr >= 0 && gr < a.nr && gz >= 0 && gz < a.nz;
        hoff[i] = ok ? H * a.sp + (int64_t)(gr + H) * a.sr + (gz + H) : -1;
    }

    auto issue = [&](int p) {
        if (p < p_end) {
            const int slot = (p - p_begin) % NB;
            const int64_t pofs = (int64_t)p * a.sp;
#pragma unroll
            for (int i = 0; i < NH; i++)
                if (hsm[i] >= 0)
                    cp_async16(tiles_s + (uint32_t)(slot * SROWS * SW + hsm[i]) * 4u,
                               a.cur + (hoff[i] >= 0 ? hoff[i] + pofs : 0), hoff[i] >= 0);
            const uint32_t d = aux_s + (uint32_t)((slot * NAUX) * 256 + tid) * 16u;
            const int64_t o = active ? own0 + pofs : 0;
            cp_async16(d, a.prev + o, active);
            cp_async16(d + 256u * 16u, a.c1 + o, active);
            cp_async16(d + 2u * 256u * 16u, a.c2 + o, active);
            if (IMG == 2) {
                cp_async16(d + 3u * 256u * 16u, a.grad + o, active);
                cp_async16(d + 4u * 256u * 16u, a.h1 + o, active);
            }
        }
        cp_async_commit();      // one group per plane, possibly empty: keeps the wait count uniform
    };

    float4 q[NQ];
#pragma unroll
    for (int i = 0; i < NQ; i++) {
        const int p = p_begin - R + i;
        q[i] = (active && p >= 0 && p < a.np) ? ld4(a.cur + own0 + (int64_t)p * a.sp) : zero4();
    }
    issue(p_begin);
    issue(p_begin + 1);

    for (int p = p_begin; p < p_end; ++p) {
        const int slot = (p - p_begin) % NB;
        float *tb = tiles + slot * SROWS * SW;
        const int64_t pofs = (int64_t)p * a.sp;
        st4(tb + (R + tr) * SW + ZH + 4 * tz, q[QC]);
        cp_async_wait<1>();         // this thread's copies for plane p have landed (plane p+1 may be in flight)
        __syncthreads();
        issue(p + 2);               // ring slot (p+2)%3 == slot of plane p-1, whose readers all passed the barrier
        float4 qn = zero4();
        const int pn = p + R + 1;
        if (p + 1 < p_end && active && pn < a.np) qn = ld4(a.cur + own0 + (int64_t)pn * a.sp);
        float4 il = zero4();
        if (a.illum && active) il = ld4(a.illum + own0 + pofs);

        if (active) {
            const float4 *ax = aux + (slot * NAUX) * 256 + tid;
            const float4 prev = ax[0], c1 = ax[256], c2 = ax[512];
            const float4 C = q[QC];
            const float4 o = point_update<R, 3, SW>(a, q, 0, tb + (R + tr) * SW + ZH + 4 * tz, prev, c1, c2, zvalid,
                                                    (a.fs && z0 <= R) ? z0 : -1);
            st4(a.out + own0 + pofs, o);
            if (IMG == 2) st4(a.grad + own0 + pofs, img4(ax[768], ax[1024], a.hist_uv ? d2u4(prev, C, o, a.inv_dt2) : C));
            if (a.illum) st4(a.illum + own0 + pofs, fma4(C, C, il));
            if (a.d2u) {
                float4 d = d2u4(prev, C, o, a.inv_dt2);
                if (zvalid < 2) d.y = 0.f;
                if (zvalid < 3) d.z = 0.f;
                if (zvalid < 4) d.w = 0.f;
                st4(a.d2u + own0 + pofs, d);
            }
        }
#pragma unroll
        for (int i = 0; i < NQ - 1; i++) q[i] = q[i + 1];
        q[NQ - 1] = qn;
    }
    cp_async_wait<0>();
}

template <int R, int IMG>
static int launch_async(const StepArgs &a, dim3 grid, cudaStream_t st)
{
    constexpr int RZ4 = (R + 3) / 4, SW = 64 + 8 * RZ4, SROWS = 16 + 2 * R, NAUX = (IMG == 2) ? 5 : 3;
    const size_t smem = (size_t)3 * SROWS * SW * 4 + (size_t)3 * NAUX * 256 * 16;
    auto kern = step3d_async_kernel<R, IMG>;
    static bool configured = false;
    if (!configured) {
        B2_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = true;
    }
    kern<<<grid, dim3(16, 16, 1), smem, st>>>(a);
    B2_CUDA(cudaGetLastError());
    count_launch();
    return 0;
}

static const int g_async = []() { const char *e = getenv("B2FWI_ASYNC_KERNEL"); return e ? atoi(e) : -1; }();   // -1: auto

// measured on 592^3 so=8 forward (fraction of HBM copy bandwidth): 0: 0.714, 1: 0.740, 2: 0.714, 3: 0.744
static const int g_tile = []() { const char *e = getenv("B2FWI_TILE"); return e ? atoi(e) : 3; }();
static void tile_shape(int ndim, int *tz, int *tr, int *blocks_per_sm)
{
    // 3-D tile variants (A/B via B2FWI_TILE): 0: 64z x 16 rows, 1: 128z x 8, 2: 64z x 32, 3: 128z x 16
    *tz = 64; *tr = 16; *blocks_per_sm = 2;
    if (ndim != 3) return;
    if (g_tile == 1) { *tz = 128; *tr = 8; }
    else if (g_tile == 2) { *tz = 64; *tr = 32; *blocks_per_sm = 1; }
    else if (g_tile == 3) { *tz = 128; *tr = 16; *blocks_per_sm = 1; }
}

template <int R, int NDIM>
static int launch_step_img(const Layout &L, const StepArgs &a, int img, dim3 grid, cudaStream_t st)
{
    if (NDIM == 3 && g_tile != 0 && img != 1 && R <= 4) {
#define B2_TILE(IMGV)                                                                                           \
        if (g_tile == 1) step_kernel<R, NDIM, IMGV, 2, 32, 8><<<grid, dim3(32, 8, 1), 0, st>>>(a);              \
        else if (g_tile == 2) step_kernel<R, NDIM, IMGV, 1, 16, 32><<<grid, dim3(16, 32, 1), 0, st>>>(a);       \
        else step_kernel<R, NDIM, IMGV, 1, 32, 16><<<grid, dim3(32, 16, 1), 0, st>>>(a);
        if (img == 0) { B2_TILE(0) } else { B2_TILE(2) }
#undef B2_TILE
        B2_CUDA(cudaGetLastError());
        count_launch();
        return 0;
    }
    dim3 block(16, 16, 1);
    switch (img) {
    case 0: step_kernel<R, NDIM, 0, 2><<<grid, block, 0, st>>>(a); break;
    case 1: step_kernel<R, NDIM, 1, 2><<<grid, block, 0, st>>>(a); break;
    case 2: step_kernel<R, NDIM, 2, 2><<<grid, block, 0, st>>>(a); break;
    default: set_error("bad imaging mode %d", img); return B2FWI_EINVAL;
    }
    B2_CUDA(cudaGetLastError());
    count_launch();
    return 0;
}

int pick_chunk(const Layout &L)
{
    // 3-D: split the streamed axis so that the grid is a whole number of waves of resident CTAs
    // (148 SMs x 2 CTAs); each chunk re-reads 2R planes of pipeline priming.
    if (L.ndim != 3) return 1;
    int tzc, trc, bps;
    tile_shape(L.ndim, &tzc, &trc, &bps);
    if (L.halo == 0 && !L.fs && tma_enabled(0)) { tma_tile_shape(L.R, &tzc, &trc); bps = 1; }     // one TMA CTA per SM
    else if (L.R > 4) { tzc = 64; trc = 16; bps = 2; }
    static int sms = 0;
    if (sms == 0) {
        int dev = 0;
        sms = 148;
        if (cudaGetDevice(&dev) == cudaSuccess)
            cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    }
    const int n_slots = bps * sms;
    const long tiles = (long)((L.nz + tzc - 1) / tzc) * ((L.nr + trc - 1) / trc);
    int best_nc = 1;
    double best_cost = 1e30;
    const int max_nc = L.np / (4 * L.R) > 0 ? L.np / (4 * L.R) : 1;
    for (int nc = 1; nc <= max_nc && nc <= 64; nc++) {
        const int chunk = (L.np + nc - 1) / nc;
        const long blocks = tiles * nc;
        const long waves = (blocks + n_slots - 1) / n_slots;
        // time ~ waves * (chunk + priming overhead counted at 1/5: only the `cur` stream is re-read)
        const double cost = (double)waves * (chunk + 2.0 * L.R / 5.0);
        if (cost < best_cost - 1e-9) { best_cost = cost; best_nc = nc; }
    }
    return (L.np + best_nc - 1) / best_nc;
}

// Kernel choice for 3-D sweeps, measured on 592^3 (fraction of the 6.46 TB/s copy bandwidth, forward sweep):
//   so <= 8 : register-staged kernel 0.71 vs cp.async pipeline 0.67;  so = 16: cp.async pipeline 0.61 vs 0.50.
// B2FWI_ASYNC_KERNEL=0/1 overrides for A/B comparison.
static bool use_async(int R) { return g_async < 0 ? R > 4 : g_async == 1; }

int launch_step(const Layout &L, StepArgs a, int img, cudaStream_t st)
{
    a.np = L.np; a.nr = L.nr; a.nz = L.nz; a.halo = L.halo; a.sp = L.sp; a.sr = L.sr; a.fs = L.fs;
    if (a.chunk <= 0) a.chunk = pick_chunk(L);
    if (tma_step_supported(L, a, img)) return launch_step_tma(L, a, img, st);
    const int nchunks = (L.ndim == 3) ? (L.np + a.chunk - 1) / a.chunk : 1;
    int tzc, trc, bps;
    tile_shape(L.ndim, &tzc, &trc, &bps);
    if (L.R > 4 || img == 1) { tzc = 64; trc = 16; }      // variants exist for so <= 8, imaging from u.dt2 / forward
    dim3 grid((L.nz + tzc - 1) / tzc, (L.nr + trc - 1) / trc, nchunks);
#define B2_CASE(r)                                                              \
    case r:                                                                     \
        if (L.ndim == 3 && img == 0 && use_async(r)) return launch_async<r, 0>(a, grid, st);   \
        if (L.ndim == 3 && img == 2 && use_async(r)) return launch_async<r, 2>(a, grid, st);   \
        return (L.ndim == 3) ? launch_step_img<r, 3>(L, a, img, grid, st)       \
                             : launch_step_img<r, 2>(L, a, img, grid, st);
    switch (L.R) {
        B2_CASE(1) B2_CASE(2) B2_CASE(3) B2_CASE(4) B2_CASE(5) B2_CASE(6) B2_CASE(7) B2_CASE(8)
    default: set_error("unsupported stencil radius %d", L.R); return B2FWI_EUNSUPPORTED;
    }
#undef B2_CASE
}

// ------------------------------------------------------------------------------------------------
// sparse points
// field[cell] += sum_j w_j * vals[pt_j] * dt^2 * vp[cell]^2, contributions in ascending point order
// (operators.py:134,221: expr = src * s**2 / m evaluated at each corner's own vp).
__global__ void inject_kernel(float *__restrict__ field, const float *__restrict__ vp, float dt,
                              const float *__restrict__ vals, int ncell,
                              const int64_t *__restrict__ cell_off, const int32_t *__restrict__ cell_ptr,
                              const int32_t *__restrict__ contrib_pt, const float *__restrict__ contrib_w,
                              float *__restrict__ d2u, const float *__restrict__ cur,
                              const float *__restrict__ prev, float inv_dt2,
                              float *__restrict__ grad, const float *__restrict__ hist)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= ncell) return;
    const int64_t off = cell_off[i];
    const float f0 = field[off];
    const float f = inject_cell(f0, cell_ptr[i], cell_ptr[i + 1], contrib_w, contrib_pt, vals, dt, vp[off]);
    field[off] = f;
    if (d2u) d2u[off] = d2u_of(prev[off], cur[off], f, inv_dt2);
    if (grad) grad[off] = img_inject_fix(grad[off], hist[off], f, f0, inv_dt2);
}

__global__ void interp_kernel(const float *__restrict__ field, float *__restrict__ out, int npoint, int ncorner,
                              const int64_t *__restrict__ corner_off, const float *__restrict__ corner_w)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= npoint) return;
    out[p] = interp_point(field, p, ncorner, corner_off, corner_w);
}

int launch_inject(float *field, const float *vp, float dt, const float *vals, const b2fwi_sparse *m,
                  float *d2u, const float *cur, const float *prev, float inv_dt2, cudaStream_t st,
                  float *grad, const float *hist)
{
    if (!m || m->ncell <= 0) return 0;
    inject_kernel<<<(m->ncell + 127) / 128, 128, 0, st>>>(field, vp, dt, vals, m->ncell, m->cell_off, m->cell_ptr,
                                                        m->contrib_pt, m->contrib_w, d2u, cur, prev, inv_dt2,
                                                        grad, hist);
    B2_CUDA(cudaGetLastError());
    count_launch();
    return 0;
}

int launch_interp(const float *field, float *out, const b2fwi_sparse *m, cudaStream_t st)
{
    if (!m || m->npoint <= 0) return 0;
    interp_kernel<<<(m->npoint + 127) / 128, 128, 0, st>>>(field, out, m->npoint, m->ncorner, m->corner_off,
                                                         m->corner_w);
    B2_CUDA(cudaGetLastError());
    count_launch();
    return 0;
}

// ------------------------------------------------------------------------------------------------
// elementwise helpers
__global__ void coeff_kernel(const float *__restrict__ vp, const float *__restrict__ damp, double dt,
                             float *__restrict__ c1, float *__restrict__ c2, int64_t n)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double v = (double)vp[i];
    float a = 0.f, b = 0.f;
    if (v > 0.0) {
        const double m = 1.0 / (v * v);
        const double den = m + dt * (double)damp[i];
        a = (float)(m / den);
        b = (float)(dt * dt / den);
    }
    c1[i] = a;
    c2[i] = b;
}

// Index box of the undamped interior. The sponge profile is a sum of non-negative 1-D profiles
// (seismic/model.py:31-49), so {c1 == 1} is a box; its extent per dimension is read off the three lines
// through the grid centre, then verified over the whole grid (box_verify_kernel) - an arbitrary damping
// field simply ends up with valid = 0 and every point reads c1.
__global__ void box_lines_kernel(const float *__restrict__ c1, int np, int nr, int nz, int64_t sp, int64_t sr,
                                 int64_t base, int *__restrict__ box)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const int n[3] = {np, nr, nz};
    const int64_t st[3] = {sp, sr, 1};
    const int c[3] = {np / 2, nr / 2, nz / 2};
    const int64_t centre = base + c[0] * sp + c[1] * sr + c[2];
    const bool ok = c1[centre] == 1.0f;
    for (int d = 0; d < 3; d++) {
        int lo = c[d], hi = c[d] + 1;
        if (ok) {
            while (lo > 0 && c1[centre + (int64_t)(lo - 1 - c[d]) * st[d]] == 1.0f) lo--;
            while (hi < n[d] && c1[centre + (int64_t)(hi - c[d]) * st[d]] == 1.0f) hi++;
        }
        box[2 * d] = lo;
        box[2 * d + 1] = ok ? hi : lo;
    }
    box[6] = ok ? 1 : 0;
    box[7] = 0;
}

__global__ void box_verify_kernel(const float *__restrict__ c1, int np, int nr, int nz, int64_t sp, int64_t sr,
                                  int64_t base, int *__restrict__ box)
{
    const int lo_p = box[0], lo_r = box[2], lo_z = box[4];
    const int64_t ep = box[1] - lo_p, er = box[3] - lo_r, ez = box[5] - lo_z;
    const int64_t total = ep * er * ez;
    bool bad = false;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int z = lo_z + (int)(i % ez), r = lo_r + (int)((i / ez) % er), p = lo_p + (int)(i / (ez * er));
        if (c1[base + p * sp + r * sr + z] != 1.0f) bad = true;
    }
    if (bad) atomicExch(box + 6, 0);
}

int launch_coeffs(const Layout &L, const float *vp, const float *damp, float dt, float *coef, cudaStream_t st)
{
    const int64_t n = L.elems;
    coeff_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(vp, damp, (double)dt, coef, coef + n, n);
    int *box = reinterpret_cast<int *>(coef + 2 * n);
    box_lines_kernel<<<1, 32, 0, st>>>(coef, L.np, L.nr, L.nz, L.sp, L.sr, L.base, box);
    box_verify_kernel<<<1184, 256, 0, st>>>(coef, L.np, L.nr, L.nz, L.sp, L.sr, L.base, box);
    B2_CUDA(cudaGetLastError());
    count_launch(3);
    return 0;
}

// Born source term: field -= (c2 * dm) * d2u   (q = -dm * u.dt2 entering U.forward with the factor c2)
__global__ void born_source_kernel(float *__restrict__ field, const float *__restrict__ c2,
                                   const float *__restrict__ dm, const float *__restrict__ d2u, int64_t n)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) field[i] = __fmaf_rn(-__fmul_rn(c2[i], dm[i]), d2u[i], field[i]);
}

int launch_born_source(const Layout &L, float *field, const float *c2, const float *dm, const float *d2u,
                       cudaStream_t st)
{
    born_source_kernel<<<(unsigned)((L.elems + 255) / 256), 256, 0, st>>>(field, c2, dm, d2u, L.elems);
    B2_CUDA(cudaGetLastError());
    count_launch();
    return 0;
}

// ------------------------------------------------------------------------------------------------
// kernel='OT4' (operators.py:38-56): H = L(u) + dt^2/12 * L((1/m) L(u)). The update is linear in H, so a step is the OT2
// sweep plus  c2 * dt^2/12 * L(vp^2 L(u))  with c2 = dt^2/(m + dt damp): two applications of a plain Laplacian kernel.
// vp^2 L(u) is taken as zero outside the padded grid (see oracle/fwi_oracle_body.inc: ot4_correction). One point per
// thread, neighbours through L1/L2: an API-completeness path (the named configurations all run OT2), not a tuned one.
template <int NDIM>
__global__ void lap_apply_kernel(const StepArgs a, int R, const float *__restrict__ f, float *__restrict__ out,
                                 const float *__restrict__ scale, float k, int mode)
{
    const int z = blockIdx.x * blockDim.x + threadIdx.x, r = blockIdx.y * blockDim.y + threadIdx.y, p = blockIdx.z;
    if (z >= a.nz || r >= a.nr) return;
    const int64_t i = (NDIM == 3 ? (int64_t)p * a.sp : 0) + (int64_t)r * a.sr + z;
    const float C = f[i];
    float lap = fmaf(a.c0, C, a.c0_lo * C);
    for (int d = 1; d <= R; d++) {
        if (NDIM == 3)
            lap = fmaf(a.cp[d], (p + d < a.np ? f[i + d * a.sp] : 0.f) + (p - d >= 0 ? f[i - d * a.sp] : 0.f), lap);
        lap = fmaf(a.cr[d], (r + d < a.nr ? f[i + d * a.sr] : 0.f) + (r - d >= 0 ? f[i - d * a.sr] : 0.f), lap);
        lap = fmaf(a.cz[d], (z + d < a.nz ? f[i + d] : 0.f) + (z - d >= 0 ? f[i - d] : 0.f), lap);
    }
    if (mode == 0) out[i] = scale[i] * scale[i] * lap;            // tmp = vp^2 L(u)
    else out[i] = fmaf(scale[i] * k, lap, out[i]);                // u+ += c2 dt^2/12 L(tmp)
}

int launch_ot4_correction(const Layout &L, const StepArgs &a0, const float *vp, float dt, float *tmp, cudaStream_t st)
{
    StepArgs a = a0;
    a.np = L.np; a.nr = L.nr; a.nz = L.nz; a.halo = L.halo; a.sp = L.sp; a.sr = L.sr;
    if (L.halo != 0 || L.fs) { set_error("OT4 needs halo 0 and no free surface"); return B2FWI_EUNSUPPORTED; }
    dim3 block(64, 4, 1), grid((L.nz + 63) / 64, (L.nr + 3) / 4, L.ndim == 3 ? L.np : 1);
    const float k = dt * dt / 12.f;
    if (L.ndim == 3) {
        lap_apply_kernel<3><<<grid, block, 0, st>>>(a, L.R, a.cur, tmp, vp, 0.f, 0);
        lap_apply_kernel<3><<<grid, block, 0, st>>>(a, L.R, tmp, a.out, a.c2, k, 1);
    } else {
        lap_apply_kernel<2><<<grid, block, 0, st>>>(a, L.R, a.cur, tmp, vp, 0.f, 0);
        lap_apply_kernel<2><<<grid, block, 0, st>>>(a, L.R, tmp, a.out, a.c2, k, 1);
    }
    B2_CUDA(cudaGetLastError());
    count_launch(2);
    return 0;
}

__global__ void accum_sq_kernel(float *__restrict__ acc, const float *__restrict__ f, int64_t n)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) acc[i] = fmaf(f[i], f[i], acc[i]);
}

int launch_accum_sq(const Layout &L, float *acc, const float *f, cudaStream_t st)
{
    accum_sq_kernel<<<(unsigned)((L.elems + 255) / 256), 256, 0, st>>>(acc, f, L.elems);
    B2_CUDA(cudaGetLastError());
    count_launch();
    return 0;
}

// fwi.py:104-129 (axes swapped: xx[i,j] = z_j, zz[i,j] = x_i)
__global__ void geometry_mask_kernel(int nx, int nz, double dx, double dz, const double *__restrict__ pts,
                                     int npts, double *__restrict__ mask)
{
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= nx * nz) return;
    const int i = idx / nz, j = idx % nz;
    const double xx = j * dz, zz = i * dx;
    const double sigma = dx + dz;
    const double inv = 1.0 / (sigma * sigma);
    double m = 1.0;
    for (int k = 0; k < npts; k++) {
        const double a = xx - pts[2 * k], b = zz - pts[2 * k + 1];
        m = m * (1.0 - exp(-.5 * (a * a + b * b) * inv));
    }
    mask[idx] = m;
}

__global__ void crop_mask_acc_kernel(int nx, int nz, int nbl, int64_t sr, int64_t base,
                                     const float *__restrict__ field, const double *__restrict__ mask,
                                     double *__restrict__ out)
{
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= nx * nz) return;
    const int i = idx / nz, j = idx % nz;
    const double f = (double)field[base + (int64_t)(i + nbl) * sr + (j + nbl)];
    out[idx] += mask ? f * mask[idx] : f;
}

int launch_geometry_mask(const b2fwi_grid *g, int nbl, const double *pts, int npts, double *mask, cudaStream_t st)
{
    const int nx = g->shape[0] - 2 * nbl, nz = g->shape[1] - 2 * nbl;
    geometry_mask_kernel<<<(nx * nz + 127) / 128, 128, 0, st>>>(nx, nz, (double)g->spacing[0], (double)g->spacing[1],
                                                              pts, npts, mask);
    B2_CUDA(cudaGetLastError());
    count_launch();
    return 0;
}

int launch_crop_mask_acc(const b2fwi_grid *g, const Layout &L, int nbl, const float *field, const double *mask,
                         double *out, cudaStream_t st)
{
    const int nx = g->shape[0] - 2 * nbl, nz = g->shape[1] - 2 * nbl;
    crop_mask_acc_kernel<<<(nx * nz + 127) / 128, 128, 0, st>>>(nx, nz, nbl, L.sr, L.base, field, mask, out);
    B2_CUDA(cudaGetLastError());
    count_launch();
    return 0;
}

}  // namespace b2fwi

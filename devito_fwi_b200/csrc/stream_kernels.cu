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
#include "common.cuh"
#include "stream_kernels.cuh"

namespace b2fwi {

static __device__ __forceinline__ float4 ld4(const float *p) { return *reinterpret_cast<const float4 *>(p); }
static __device__ __forceinline__ float4 ldg4(const float *p) { return __ldg(reinterpret_cast<const float4 *>(p)); }
static __device__ __forceinline__ void st4(float *p, float4 v) { *reinterpret_cast<float4 *>(p) = v; }
static __device__ __forceinline__ float4 zero4() { return make_float4(0.f, 0.f, 0.f, 0.f); }
// u.dt2 from three time levels; one fixed operation order everywhere so that a checkpointed
// gradient (u.dt2 stored by the recompute sweep) is bitwise identical to the full-history one.
static __device__ __forceinline__ float d2u_of(float um, float uc, float up, float inv_dt2)
{
    return __fmul_rn(__fadd_rn(__fmaf_rn(-2.f, uc, um), up), inv_dt2);
}

template <int R, int NDIM, int IMG>
__global__ void __launch_bounds__(256, 2) step_kernel(const __grid_constant__ StepArgs a)
{
    constexpr int TR = 16;                 // rows per tile
    constexpr int RZ4 = (R + 3) / 4;       // z-halo width in float4
    constexpr int ZH = 4 * RZ4;            // z-halo width in floats
    constexpr int SW = 64 + 2 * ZH;        // shared row width (floats)
    constexpr int SROWS = TR + 2 * R;
    constexpr int NHALO = 2 * R * 16 + TR * 2 * RZ4;
    constexpr int NH = (NHALO + 255) / 256;
    constexpr int NQ = (NDIM == 3) ? 2 * R + 1 : 1;
    constexpr int QC = (NDIM == 3) ? R : 0;

    __shared__ __align__(16) float tile[2][SROWS][SW];

    const int tz = threadIdx.x, tr = threadIdx.y, tid = tr * 16 + tz;
    const int ztile0 = blockIdx.x * 64;
    const int z0 = ztile0 + tz * 4;
    const int r0 = blockIdx.y * TR;
    const int r = r0 + tr;
    const bool active = (r < a.nr) && (z0 < a.nz);
    const int p_begin = (NDIM == 3) ? (int)blockIdx.z * a.chunk : 0;
    const int p_end = (NDIM == 3) ? min(p_begin + a.chunk, a.np) : 1;
    const int64_t H = a.halo;
    // offset of this thread's float4 in plane 0
    const int64_t own0 = (NDIM == 3 ? H * a.sp : 0) + (int64_t)(r + H) * a.sr + (z0 + H);

    // ---- halo work items of this thread (fixed across planes)
    int64_t hoff[NH];      // global offset inside plane 0, or -1
    int hsm[NH];           // shared offset (floats) inside one buffer, or -1
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
        hoff[i] = ok ? (NDIM == 3 ? H * a.sp : 0) + (int64_t)(gr + H) * a.sr + (gz + H) : -1;
    }

    // ---- prologue: register pipeline along the plane axis, halo of the first plane
    float4 q[NQ];
    if (NDIM == 3) {
#pragma unroll
        for (int i = 0; i < NQ; i++) {
            const int p = p_begin - R + i;
            q[i] = (active && p >= 0 && p < a.np) ? ld4(a.cur + own0 + (int64_t)p * a.sp) : zero4();
        }
    } else {
        q[0] = active ? ld4(a.cur + own0) : zero4();
    }
    float4 hreg[NH];
#pragma unroll
    for (int i = 0; i < NH; i++)
        hreg[i] = (hoff[i] >= 0) ? ld4(a.cur + hoff[i] + (int64_t)p_begin * (NDIM == 3 ? a.sp : 0)) : zero4();

    for (int p = p_begin; p < p_end; ++p) {
        const int buf = (p - p_begin) & 1;
        float *tb = &tile[buf][0][0];
        const int64_t pofs = (NDIM == 3) ? (int64_t)p * a.sp : 0;
        // stage plane p
        st4(tb + (R + tr) * SW + ZH + 4 * tz, q[QC]);
#pragma unroll
        for (int i = 0; i < NH; i++)
            if (hsm[i] >= 0) st4(tb + hsm[i], hreg[i]);

        // pointwise operands of plane p
        float4 prev = zero4(), c1 = zero4(), c2 = zero4();
        float4 g4 = zero4(), h0 = zero4(), h1 = zero4(), h2 = zero4(), il = zero4();
        if (active) {
            prev = ld4(a.prev + own0 + pofs);
            c1 = ldg4(a.c1 + own0 + pofs);
            c2 = ldg4(a.c2 + own0 + pofs);
            if (IMG != 0) {
                g4 = ld4(a.grad + own0 + pofs);
                h1 = ldg4(a.h1 + own0 + pofs);
                if (IMG == 1) {
                    h0 = ldg4(a.h0 + own0 + pofs);
                    h2 = ldg4(a.h2 + own0 + pofs);
                }
            }
            if (a.illum) il = ld4(a.illum + own0 + pofs);
        }
        // prefetch for plane p+1
        float4 qn = zero4();
        if (p + 1 < p_end) {
            if (NDIM == 3) {
                const int pn = p + R + 1;
                if (active && pn < a.np) qn = ld4(a.cur + own0 + (int64_t)pn * a.sp);
            }
#pragma unroll
            for (int i = 0; i < NH; i++)
                hreg[i] = (hoff[i] >= 0) ? ld4(a.cur + hoff[i] + pofs + a.sp) : zero4();
        }
        __syncthreads();

        if (active) {
            const float4 C = q[QC];
            float lap[4] = {fmaf(a.c0, C.x, a.c0_lo * C.x), fmaf(a.c0, C.y, a.c0_lo * C.y),
                            fmaf(a.c0, C.z, a.c0_lo * C.z), fmaf(a.c0, C.w, a.c0_lo * C.w)};
            if (NDIM == 3) {
#pragma unroll
                for (int k = 1; k <= R; k++) {
                    const float4 A = q[QC + k], B = q[QC - k];
                    const float c = a.cp[k];
                    lap[0] = fmaf(c, A.x + B.x, lap[0]);
                    lap[1] = fmaf(c, A.y + B.y, lap[1]);
                    lap[2] = fmaf(c, A.z + B.z, lap[2]);
                    lap[3] = fmaf(c, A.w + B.w, lap[3]);
                }
            }
            const float *ctr = tb + (R + tr) * SW + ZH + 4 * tz;
#pragma unroll
            for (int k = 1; k <= R; k++) {
                const float4 A = ld4(ctr + k * SW), B = ld4(ctr - k * SW);
                const float c = a.cr[k];
                lap[0] = fmaf(c, A.x + B.x, lap[0]);
                lap[1] = fmaf(c, A.y + B.y, lap[1]);
                lap[2] = fmaf(c, A.z + B.z, lap[2]);
                lap[3] = fmaf(c, A.w + B.w, lap[3]);
            }
            float zl[ZH + 4 + ZH];
#pragma unroll
            for (int i = 0; i < RZ4; i++) {
                const float4 Lq = ld4(ctr - ZH + 4 * i), Rq = ld4(ctr + 4 + 4 * i);
                zl[4 * i + 0] = Lq.x; zl[4 * i + 1] = Lq.y; zl[4 * i + 2] = Lq.z; zl[4 * i + 3] = Lq.w;
                zl[ZH + 4 + 4 * i + 0] = Rq.x; zl[ZH + 4 + 4 * i + 1] = Rq.y;
                zl[ZH + 4 + 4 * i + 2] = Rq.z; zl[ZH + 4 + 4 * i + 3] = Rq.w;
            }
            zl[ZH + 0] = C.x; zl[ZH + 1] = C.y; zl[ZH + 2] = C.z; zl[ZH + 3] = C.w;
#pragma unroll
            for (int j = 0; j < 4; j++)
#pragma unroll
                for (int k = 1; k <= R; k++)
                    lap[j] = fmaf(a.cz[k], zl[ZH + j + k] + zl[ZH + j - k], lap[j]);

            float o[4];
            const float Cv[4] = {C.x, C.y, C.z, C.w};
            const float Pv[4] = {prev.x, prev.y, prev.z, prev.w};
            const float c1v[4] = {c1.x, c1.y, c1.z, c1.w};
            const float c2v[4] = {c2.x, c2.y, c2.z, c2.w};
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const float t = fmaf(c1v[j], Cv[j] - Pv[j], Cv[j]);
                o[j] = (z0 + j < a.nz) ? fmaf(c2v[j], lap[j], t) : 0.f;
            }
            st4(a.out + own0 + pofs, make_float4(o[0], o[1], o[2], o[3]));

            if (IMG != 0) {
                float d2[4];
                if (IMG == 1) {
                    d2[0] = d2u_of(h0.x, h1.x, h2.x, a.inv_dt2);
                    d2[1] = d2u_of(h0.y, h1.y, h2.y, a.inv_dt2);
                    d2[2] = d2u_of(h0.z, h1.z, h2.z, a.inv_dt2);
                    d2[3] = d2u_of(h0.w, h1.w, h2.w, a.inv_dt2);
                } else {
                    d2[0] = h1.x; d2[1] = h1.y; d2[2] = h1.z; d2[3] = h1.w;
                }
                g4.x = __fmaf_rn(-d2[0], C.x, g4.x);
                g4.y = __fmaf_rn(-d2[1], C.y, g4.y);
                g4.z = __fmaf_rn(-d2[2], C.z, g4.z);
                g4.w = __fmaf_rn(-d2[3], C.w, g4.w);
                st4(a.grad + own0 + pofs, g4);
            }
            if (a.illum) {
                il.x = fmaf(C.x, C.x, il.x); il.y = fmaf(C.y, C.y, il.y);
                il.z = fmaf(C.z, C.z, il.z); il.w = fmaf(C.w, C.w, il.w);
                st4(a.illum + own0 + pofs, il);
            }
            if (a.d2u) {
                float4 d;
                d.x = (z0 + 0 < a.nz) ? d2u_of(prev.x, C.x, o[0], a.inv_dt2) : 0.f;
                d.y = (z0 + 1 < a.nz) ? d2u_of(prev.y, C.y, o[1], a.inv_dt2) : 0.f;
                d.z = (z0 + 2 < a.nz) ? d2u_of(prev.z, C.z, o[2], a.inv_dt2) : 0.f;
                d.w = (z0 + 3 < a.nz) ? d2u_of(prev.w, C.w, o[3], a.inv_dt2) : 0.f;
                st4(a.d2u + own0 + pofs, d);
            }
        }
        if (NDIM == 3) {
#pragma unroll
            for (int i = 0; i < NQ - 1; i++) q[i] = q[i + 1];
            q[NQ - 1] = qn;
        }
    }
}

template <int R, int NDIM>
static int launch_step_img(const Layout &L, const StepArgs &a, int img, dim3 grid, cudaStream_t st)
{
    dim3 block(16, 16, 1);
    switch (img) {
    case 0: step_kernel<R, NDIM, 0><<<grid, block, 0, st>>>(a); break;
    case 1: step_kernel<R, NDIM, 1><<<grid, block, 0, st>>>(a); break;
    case 2: step_kernel<R, NDIM, 2><<<grid, block, 0, st>>>(a); break;
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
    static int n_slots = 0;
    if (n_slots == 0) {
        int dev = 0, sms = 148;
        if (cudaGetDevice(&dev) == cudaSuccess)
            cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        n_slots = 2 * sms;
    }
    const long tiles = (long)((L.nz + 63) / 64) * ((L.nr + 15) / 16);
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

int launch_step(const Layout &L, StepArgs a, int img, cudaStream_t st)
{
    a.np = L.np; a.nr = L.nr; a.nz = L.nz; a.halo = L.halo; a.sp = L.sp; a.sr = L.sr;
    if (a.chunk <= 0) a.chunk = pick_chunk(L);
    const int nchunks = (L.ndim == 3) ? (L.np + a.chunk - 1) / a.chunk : 1;
    dim3 grid((L.nz + 63) / 64, (L.nr + 15) / 16, nchunks);
#define B2_CASE(r)                                                              \
    case r:                                                                     \
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
                              const float *__restrict__ prev, float inv_dt2)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= ncell) return;
    const int64_t off = cell_off[i];
    const float v = vp[off];
    float f = field[off];
    for (int j = cell_ptr[i]; j < cell_ptr[i + 1]; j++)
        f += contrib_w[j] * vals[contrib_pt[j]] * dt * dt * v * v;
    field[off] = f;
    if (d2u) d2u[off] = d2u_of(prev[off], cur[off], f, inv_dt2);
}

// out[p] = sum_c w_c * field[c]  (operators.py:137,176)
__global__ void interp_kernel(const float *__restrict__ field, float *__restrict__ out, int npoint, int ncorner,
                              const int64_t *__restrict__ corner_off, const float *__restrict__ corner_w)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= npoint) return;
    float sum = 0.f;
    for (int c = 0; c < ncorner; c++) {
        const int64_t off = corner_off[(int64_t)p * ncorner + c];
        if (off >= 0) sum += corner_w[(int64_t)p * ncorner + c] * field[off];
    }
    out[p] = sum;
}

int launch_inject(float *field, const float *vp, float dt, const float *vals, const b2fwi_sparse *m,
                  float *d2u, const float *cur, const float *prev, float inv_dt2, cudaStream_t st)
{
    if (!m || m->ncell <= 0) return 0;
    inject_kernel<<<(m->ncell + 127) / 128, 128, 0, st>>>(field, vp, dt, vals, m->ncell, m->cell_off, m->cell_ptr,
                                                        m->contrib_pt, m->contrib_w, d2u, cur, prev, inv_dt2);
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

int launch_coeffs(const Layout &L, const float *vp, const float *damp, float dt, float *coef, cudaStream_t st)
{
    const int64_t n = L.elems;
    coeff_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(vp, damp, (double)dt, coef, coef + n, n);
    B2_CUDA(cudaGetLastError());
    count_launch();
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

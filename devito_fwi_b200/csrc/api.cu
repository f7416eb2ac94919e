// api.cu -- extern "C" entry points of libb2fwi.so (see include/b2fwi.h) and the host-side
// time loops of the streaming engine.
#include <stdarg.h>
#include <atomic>
#include <string.h>

#include "common.cuh"
#include "stream_kernels.cuh"

namespace b2fwi {

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

void set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

void laplace_coeffs(int R, double *c)
{
    // c_k = 2 (-1)^(k+1) (R!)^2 / (k^2 (R-k)! (R+k)!),  c_0 = -2 sum_k c_k   (field.laplace, operators.py:56)
    double s = 0.0;
    for (int k = 1; k <= R; k++) {
        double r = 1.0;
        for (int j = 1; j <= k; j++) r *= (double)(R - k + j) / (double)(R + j);
        c[k] = ((k & 1) ? 2.0 : -2.0) * r / ((double)k * (double)k);
        s += c[k];
    }
    c[0] = -2.0 * s;
}

int make_layout(const b2fwi_grid *g, Layout *L)
{
    B2_CHECK_ARG(g != nullptr, "grid is NULL");
    B2_CHECK_ARG(g->ndim == 2 || g->ndim == 3, "ndim must be 2 or 3 (got %d)", g->ndim);
    B2_CHECK_ARG(g->space_order >= 2 && g->space_order <= 2 * B2FWI_MAX_R && (g->space_order % 2) == 0,
                 "space_order must be even in [2, 16] (got %d)", g->space_order);
    const int R = g->space_order / 2;
    B2_CHECK_ARG(g->halo >= 0 && g->halo % 4 == 0, "halo must be a non-negative multiple of 4 (got %d)", g->halo);
    for (int d = 0; d < g->ndim; d++) {
        B2_CHECK_ARG(g->shape[d] >= 1, "shape[%d] = %d", d, g->shape[d]);
        B2_CHECK_ARG(g->spacing[d] > 0.f, "spacing[%d] = %g", d, (double)g->spacing[d]);
    }
    const int64_t H = g->halo;
    L->ndim = g->ndim;
    L->halo = g->halo;
    L->R = R;
    L->fs = g->fs ? 1 : 0;
    L->ot4 = (g->kernel == 1) ? 1 : 0;
    B2_CHECK_ARG(g->kernel == 0 || g->kernel == 1, "kernel must be 0 (OT2) or 1 (OT4)");
    B2_CHECK_ARG(!(L->ot4 && L->fs), "kernel='OT4' with a free surface is not supported");
    if (g->ndim == 2) {
        L->np = 1; L->nr = g->shape[0]; L->nz = g->shape[1];
        L->sr = ((int64_t)L->nz + 2 * H + 31) / 32 * 32;
        L->sp = 0;
        L->elems = (L->nr + 2 * H) * L->sr;
        L->base = H * L->sr + H;
        L->inv_h2[0] = 0.0;
        L->inv_h2[1] = 1.0 / ((double)g->spacing[0] * (double)g->spacing[0]);
        L->inv_h2[2] = 1.0 / ((double)g->spacing[1] * (double)g->spacing[1]);
    } else {
        L->np = g->shape[0]; L->nr = g->shape[1]; L->nz = g->shape[2];
        L->sr = ((int64_t)L->nz + 2 * H + 31) / 32 * 32;
        L->sp = (L->nr + 2 * H) * L->sr;
        L->elems = (L->np + 2 * H) * L->sp;
        L->base = H * L->sp + H * L->sr + H;
        for (int d = 0; d < 3; d++) L->inv_h2[d] = 1.0 / ((double)g->spacing[d] * (double)g->spacing[d]);
    }
    return 0;
}

void fill_stencil_weights(const Layout &L, StepArgs *a)
{
    double c[B2FWI_MAX_R + 1];
    laplace_coeffs(L.R, c);
    for (int k = 0; k <= B2FWI_MAX_R; k++) a->cp[k] = a->cr[k] = a->cz[k] = 0.f;
    // The centre weight must cancel the ROUNDED side weights exactly: a mismatch of one fp32 ulp
    // acts as a spurious mass term eps*u/h^2 whose phase error grows linearly in time and was
    // measured at 2.7e-5 relative L2 on Marmousi traces (vs 4.7e-6 with the exact cancellation).
    // It is therefore carried as an unevaluated hi + lo pair (one extra FMA per point).
    double centre = 0.0;
    for (int k = 1; k <= L.R; k++) {
        a->cp[k] = (float)(c[k] * L.inv_h2[0]);
        a->cr[k] = (float)(c[k] * L.inv_h2[1]);
        a->cz[k] = (float)(c[k] * L.inv_h2[2]);
        centre -= 2.0 * ((double)a->cr[k] + (double)a->cz[k] + (L.ndim == 3 ? (double)a->cp[k] : 0.0));
    }
    a->c0 = (float)centre;
    a->c0_lo = (float)(centre - (double)a->c0);
}

// scratch slice for the OT4 correction (stream-ordered allocation, zero in the pitch padding)
struct Ot4Scratch {
    float *p = nullptr;
    cudaStream_t st;
    int get(const Layout &L, cudaStream_t s)
    {
        st = s;
        if (!L.ot4) return 0;
        B2_CUDA(cudaMallocAsync(reinterpret_cast<void **>(&p), (size_t)L.elems * sizeof(float), s));
        B2_CUDA(cudaMemsetAsync(p, 0, (size_t)L.elems * sizeof(float), s));
        return 0;
    }
    ~Ot4Scratch() { if (p) cudaFreeAsync(p, st); }
};

static int check_time(int nt, int time_m, int time_M)
{
    B2_CHECK_ARG(nt >= 3, "nt must be >= 3 (got %d)", nt);
    B2_CHECK_ARG(time_m >= 1 && time_M <= nt - 2, "time range [%d, %d] outside [1, nt-2] with nt=%d", time_m, time_M,
                 nt);
    return 0;
}

}  // namespace b2fwi

using namespace b2fwi;

extern "C" {

int32_t b2fwi_version(void) { return B2FWI_VERSION; }

const char *b2fwi_last_error(void) { return g_err; }

int64_t b2fwi_launch_count(void) { return (int64_t)g_launches.load(); }

int b2fwi_set_option(const char *name, int32_t value)
{
    B2_CHECK_ARG(name != nullptr, "option name is NULL");
    if (strcmp(name, "tma") == 0) {
        const int old = get_tma_mask();
        set_tma_mask(value);
        return old;
    }
    if (strcmp(name, "fuse") == 0) {
        const int old = get_fuse();
        set_fuse(value);
        return old;
    }
    set_error("unknown option '%s'", name);
    return B2FWI_EINVAL;
}

int b2fwi_field_layout(const b2fwi_grid *g, int64_t stride_out[3], int64_t *base_out, int64_t *elems_out)
{
    Layout L;
    int rc = make_layout(g, &L);
    if (rc) return rc;
    if (stride_out) {
        if (g->ndim == 2) { stride_out[0] = L.sr; stride_out[1] = 1; stride_out[2] = 0; }
        else { stride_out[0] = L.sp; stride_out[1] = L.sr; stride_out[2] = 1; }
    }
    if (base_out) *base_out = L.base;
    if (elems_out) *elems_out = L.elems;
    return 0;
}

int b2fwi_prepare_coeffs(const b2fwi_grid *g, const float *vp, const float *damp, float dt, float *coef,
                         void *stream)
{
    Layout L;
    int rc = make_layout(g, &L);
    if (rc) return rc;
    B2_CHECK_ARG(vp && damp && coef, "NULL field pointer");
    B2_CHECK_ARG(dt > 0.f, "dt must be positive");
    return launch_coeffs(L, vp, damp, dt, coef, (cudaStream_t)stream);
}

int b2fwi_forward(const b2fwi_grid *g, const float *vp, const float *coef, float dt,
                  int32_t nt, int32_t time_m, int32_t time_M,
                  const float *src, const b2fwi_sparse *src_map,
                  float *rec, const b2fwi_sparse *rec_map,
                  float *u, int32_t save, float *illum, float *d2u_out, int32_t d2u_t0, void *stream)
{
    Layout L;
    int rc = make_layout(g, &L);
    if (rc) return rc;
    if ((rc = check_time(nt, time_m, time_M))) return rc;
    B2_CHECK_ARG(vp && coef && u, "NULL field pointer");
    const int nsrc = src_map ? src_map->npoint : 0, nrec = rec_map ? rec_map->npoint : 0;
    B2_CHECK_ARG(nsrc == 0 || src, "src is NULL with %d source points", nsrc);
    B2_CHECK_ARG(nrec == 0 || rec, "rec is NULL with %d receiver points", nrec);
    cudaStream_t st = (cudaStream_t)stream;
    StepArgs a;
    memset(&a, 0, sizeof(a));
    fill_stencil_weights(L, &a);
    a.c1 = coef; a.c2 = coef + L.elems;
    a.box = reinterpret_cast<const int *>(coef + 2 * L.elems);
    a.inv_dt2 = 1.f / (dt * dt);
    a.chunk = pick_chunk(L);
    Ot4Scratch ot4;
    if ((rc = ot4.get(L, st))) return rc;
    B2_CHECK_ARG(!(L.ot4 && d2u_out), "u.dt2 output is not available with kernel='OT4'");
    for (int time = time_m; time <= time_M; time++) {
        const int64_t sn = save ? time + 1 : (time + 1) % 3, sc = save ? time : time % 3,
                      sp = save ? time - 1 : (time - 1) % 3;
        float *un = u + sn * L.elems;
        const float *uc = u + sc * L.elems, *up = u + sp * L.elems;
        a.out = un; a.cur = uc; a.prev = up;
        a.illum = illum;
        a.d2u = d2u_out ? d2u_out + (int64_t)(time - d2u_t0) * L.elems : nullptr;
        // injection / interpolation inside the sweep kernel where it can take them (TMA sweeps with tables), else
        // as their own launches after it
        const bool tma = tma_step_supported(L, a, 0);
        const bool f_inj = tma && nsrc > 0 && a.chunk <= 1024 && tma_fusable(L, src_map, 1), f_itp = tma && nrec > 0 && tma_fusable(L, rec_map, 2);
        a.inj.row_tile = a.itp.row_tile = 0;
        if (f_inj) { a.inj = *src_map; a.inj_vals = src + (int64_t)time * nsrc; a.vp = vp; a.dt = dt; }
        if (f_itp) { a.itp = *rec_map; a.itp_out = rec + (int64_t)time * nrec; }
        if ((rc = launch_step(L, a, 0, st))) return rc;
        if (L.ot4 && (rc = launch_ot4_correction(L, a, vp, dt, ot4.p, st))) return rc;
        if (nsrc > 0 && !f_inj &&
            (rc = launch_inject(un, vp, dt, src + (int64_t)time * nsrc, src_map, a.d2u, uc, up, a.inv_dt2, st)))
            return rc;
        if (nrec > 0 && !f_itp && (rc = launch_interp(uc, rec + (int64_t)time * nrec, rec_map, st))) return rc;
    }
    if (illum && time_m <= time_M && time_M == nt - 2) {     // the call producing the last slice adds it
        const int64_t sl = save ? time_M + 1 : (time_M + 1) % 3;
        if ((rc = launch_accum_sq(L, illum, u + sl * L.elems, st))) return rc;
    }
    return 0;
}

static int backward(const b2fwi_grid *g, const float *vp, const float *coef, float dt,
                    int32_t nt, int32_t time_m, int32_t time_M,
                    const float *rec, const b2fwi_sparse *rec_map,
                    float *srca, const b2fwi_sparse *src_map,
                    const float *hist, int32_t hist_kind, int32_t hist_t0,
                    float *v, float *grad, void *stream)
{
    Layout L;
    int rc = make_layout(g, &L);
    if (rc) return rc;
    if ((rc = check_time(nt, time_m, time_M))) return rc;
    B2_CHECK_ARG(vp && coef && v, "NULL field pointer");
    const int nrec = rec_map ? rec_map->npoint : 0, nsrc = src_map ? src_map->npoint : 0;
    B2_CHECK_ARG(nrec == 0 || rec, "rec is NULL with %d receiver points", nrec);
    int img = 0;
    if (grad) {
        B2_CHECK_ARG(hist != nullptr, "hist is NULL");
        B2_CHECK_ARG(hist_kind == B2FWI_HIST_U || hist_kind == B2FWI_HIST_D2U || hist_kind == B2FWI_HIST_UVDT2,
                     "bad hist_kind %d", hist_kind);
        B2_CHECK_ARG(!(L.ot4 && hist_kind == B2FWI_HIST_UVDT2), "kernel='OT4' images from the saved wavefield (B2FWI_HIST_U)");
        img = (hist_kind == B2FWI_HIST_UVDT2) ? B2FWI_HIST_D2U : hist_kind;
    }
    cudaStream_t st = (cudaStream_t)stream;
    StepArgs a;
    memset(&a, 0, sizeof(a));
    fill_stencil_weights(L, &a);
    a.c1 = coef; a.c2 = coef + L.elems;
    a.box = reinterpret_cast<const int *>(coef + 2 * L.elems);
    a.inv_dt2 = 1.f / (dt * dt);
    a.chunk = pick_chunk(L);
    a.grad = grad;
    a.hist_uv = (grad && hist_kind == B2FWI_HIST_UVDT2) ? 1 : 0;
    Ot4Scratch ot4;
    if ((rc = ot4.get(L, st))) return rc;
    for (int time = time_M; time >= time_m; time--) {
        float *vn = v + (int64_t)((time - 1) % 3) * L.elems;
        const float *vc = v + (int64_t)(time % 3) * L.elems, *vq = v + (int64_t)((time + 1) % 3) * L.elems;
        a.out = vn; a.cur = vc; a.prev = vq;
        if (img) {
            const float *h = hist + (int64_t)(time - hist_t0) * L.elems;
            a.h1 = h;
            a.h0 = (img == B2FWI_HIST_U) ? h - L.elems : nullptr;
            a.h2 = (img == B2FWI_HIST_U) ? h + L.elems : nullptr;
        }
        const bool tma = tma_step_supported(L, a, img);
        const bool f_inj = tma && nrec > 0 && a.chunk <= 1024 && tma_fusable(L, rec_map, 1);
        const bool f_itp = tma && srca && nsrc > 0 && tma_fusable(L, src_map, 2);
        a.inj.row_tile = a.itp.row_tile = 0;
        if (f_inj) { a.inj = *rec_map; a.inj_vals = rec + (int64_t)time * nrec; a.vp = vp; a.dt = dt; }
        if (f_itp) { a.itp = *src_map; a.itp_out = srca + (int64_t)time * nsrc; }
        if ((rc = launch_step(L, a, img, st))) return rc;
        if (L.ot4 && (rc = launch_ot4_correction(L, a, vp, dt, ot4.p, st))) return rc;
        if (nrec > 0 && !f_inj &&
            (rc = launch_inject(vn, vp, dt, rec + (int64_t)time * nrec, rec_map, nullptr, nullptr, nullptr, a.inv_dt2, st,
                                a.hist_uv ? grad : nullptr, a.hist_uv ? a.h1 : nullptr)))
            return rc;
        if (srca && nsrc > 0 && !f_itp && (rc = launch_interp(vc, srca + (int64_t)time * nsrc, src_map, st))) return rc;
    }
    return 0;
}

int b2fwi_gradient(const b2fwi_grid *g, const float *vp, const float *coef, float dt,
                   int32_t nt, int32_t time_m, int32_t time_M,
                   const float *rec, const b2fwi_sparse *rec_map,
                   const float *hist, int32_t hist_kind, int32_t hist_t0,
                   float *v, float *grad, void *stream)
{
    B2_CHECK_ARG(grad != nullptr, "grad is NULL");
    return backward(g, vp, coef, dt, nt, time_m, time_M, rec, rec_map, nullptr, nullptr, hist, hist_kind, hist_t0, v,
                    grad, stream);
}

int b2fwi_adjoint(const b2fwi_grid *g, const float *vp, const float *coef, float dt,
                  int32_t nt, int32_t time_m, int32_t time_M,
                  const float *rec, const b2fwi_sparse *rec_map,
                  float *srca, const b2fwi_sparse *src_map,
                  float *v, void *stream)
{
    return backward(g, vp, coef, dt, nt, time_m, time_M, rec, rec_map, srca, src_map, nullptr, 0, 0, v, nullptr,
                    stream);
}

int b2fwi_born(const b2fwi_grid *g, const float *vp, const float *coef, float dt,
               int32_t nt, int32_t time_m, int32_t time_M,
               const float *src, const b2fwi_sparse *src_map,
               float *rec, const b2fwi_sparse *rec_map,
               const float *dm, float *u, float *U, float *scratch, void *stream)
{
    Layout L;
    int rc = make_layout(g, &L);
    if (rc) return rc;
    if ((rc = check_time(nt, time_m, time_M))) return rc;
    B2_CHECK_ARG(vp && coef && u && U && dm && scratch, "NULL field pointer");
    B2_CHECK_ARG(!L.ot4, "the Born operator is implemented for kernel='OT2' only");
    const int nsrc = src_map ? src_map->npoint : 0, nrec = rec_map ? rec_map->npoint : 0;
    B2_CHECK_ARG(nsrc == 0 || src, "src is NULL with %d source points", nsrc);
    B2_CHECK_ARG(nrec == 0 || rec, "rec is NULL with %d receiver points", nrec);
    cudaStream_t st = (cudaStream_t)stream;
    StepArgs a;
    memset(&a, 0, sizeof(a));
    fill_stencil_weights(L, &a);
    a.c1 = coef; a.c2 = coef + L.elems;
    a.box = reinterpret_cast<const int *>(coef + 2 * L.elems);
    a.inv_dt2 = 1.f / (dt * dt);
    a.chunk = pick_chunk(L);
    for (int time = time_m; time <= time_M; time++) {
        const int64_t sn = (time + 1) % 3, sc = time % 3, sp = (time - 1) % 3;
        // background wavefield, u.dt2[time] into the scratch slice (the injection kernel patches it at the source cells)
        a.out = u + sn * L.elems; a.cur = u + sc * L.elems; a.prev = u + sp * L.elems;
        a.d2u = scratch;
        if ((rc = launch_step(L, a, 0, st))) return rc;
        if (nsrc > 0 && (rc = launch_inject(u + sn * L.elems, vp, dt, src + (int64_t)time * nsrc, src_map, scratch,
                                            a.cur, a.prev, a.inv_dt2, st)))
            return rc;
        // scattered wavefield
        a.out = U + sn * L.elems; a.cur = U + sc * L.elems; a.prev = U + sp * L.elems;
        a.d2u = nullptr;
        if ((rc = launch_step(L, a, 0, st))) return rc;
        if ((rc = launch_born_source(L, U + sn * L.elems, a.c2, dm, scratch, st))) return rc;
        if (nrec > 0 && (rc = launch_interp(U + sc * L.elems, rec + (int64_t)time * nrec, rec_map, st))) return rc;
    }
    return 0;
}

int b2fwi_geometry_mask(const b2fwi_grid *g, int32_t nbl, const double *pts, int32_t npts, double *mask_out,
                        void *stream)
{
    Layout L;
    int rc = make_layout(g, &L);
    if (rc) return rc;
    B2_CHECK_ARG(g->ndim == 2, "geometry mask is defined for 2-D models only (fwi.py:104-129)");
    B2_CHECK_ARG(nbl >= 0 && g->shape[0] > 2 * nbl && g->shape[1] > 2 * nbl, "bad nbl %d", nbl);
    B2_CHECK_ARG(pts && mask_out && npts >= 0, "NULL pointer");
    return launch_geometry_mask(g, nbl, pts, npts, mask_out, (cudaStream_t)stream);
}

int b2fwi_crop_mask_accumulate(const b2fwi_grid *g, int32_t nbl, const float *field, const double *mask,
                               double *out, void *stream)
{
    Layout L;
    int rc = make_layout(g, &L);
    if (rc) return rc;
    B2_CHECK_ARG(g->ndim == 2, "crop/mask accumulate is defined for 2-D models only");
    B2_CHECK_ARG(nbl >= 0 && g->shape[0] > 2 * nbl && g->shape[1] > 2 * nbl, "bad nbl %d", nbl);
    B2_CHECK_ARG(field && out, "NULL pointer");
    return launch_crop_mask_acc(g, L, nbl, field, mask, out, (cudaStream_t)stream);
}

}  // extern "C"

// res2d_api.cu -- extern "C" entry points of the SM-resident 2-D engine, the decomposition planner,
// and small device helpers of the fwi.py objective (window accumulate, L2 misfit).
#include <string.h>

#include "common.cuh"
#include "resident2d.cuh"
#include "stream_kernels.cuh"

namespace b2fwi {

int launch_res2d(const Res2dArgs &a, int R, int P, int mode, cudaStream_t st);
size_t res2d_smem_bytes(const Res2dArgs &a, int P);
int res2d_max_clusters(const Res2dArgs &a, int R, int P, int *out);

static const int kMaxSmem = 232448;   // 227 KB opt-in limit per CTA on sm_100

static int max_threads_for(int P) { return P <= 8 ? 480 : P <= 12 ? 384 : 288; }

__global__ void bcoef_kernel(const float *__restrict__ vp, double dt, float *__restrict__ B, int64_t n)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        const double v = (double)vp[i];
        B[i] = (float)(dt * dt * v * v);
    }
}

__global__ void window_mask_acc_kernel(int nx, int nz, const float *__restrict__ field, int64_t row_stride, int col0,
                                       const double *__restrict__ mask, double *__restrict__ out)
{
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= nx * nz) return;
    const int i = idx / nz, j = idx - i * nz;
    const double f = (double)field[(int64_t)i * row_stride + col0 + j];
    out[idx] += mask ? f * mask[idx] : f;
}

// out[i,j] += sum_s field[s][i][col0+j] * mask[s][i][j], shots in ascending order (deterministic)
__global__ void window_mask_acc_batch_kernel(int nshots, int nx, int nz, const float *__restrict__ field,
                                             int64_t shot_stride, int64_t row_stride, int col0,
                                             const double *__restrict__ mask, double *__restrict__ out)
{
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= nx * nz) return;
    const int i = idx / nz, j = idx - i * nz;
    double acc = out[idx];
    for (int s = 0; s < nshots; s++) {
        const double f = (double)field[s * shot_stride + (int64_t)i * row_stride + col0 + j];
        acc += mask ? f * mask[(int64_t)s * nx * nz + idx] : f;
    }
    out[idx] = acc;
}

// stage 1: per-block partial sums (fixed order), stage 2: one block adds them up in index order
__global__ void l2_stage1(const float *__restrict__ syn, const float *__restrict__ obs, const float *__restrict__ dw,
                          int64_t n, float *__restrict__ res, double *__restrict__ partial)
{
    __shared__ double sh[256];
    double acc = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        float r;
        if (dw) r = (syn[i] - dw[i]) - (obs[i] - dw[i]);
        else r = syn[i] - obs[i];
        res[i] = r;
        acc += (double)r * (double)r;
    }
    sh[threadIdx.x] = acc;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if (threadIdx.x < s) sh[threadIdx.x] += sh[threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0) partial[blockIdx.x] = sh[0];
}

__global__ void l2_stage2(const double *__restrict__ partial, int nblocks, double *__restrict__ fval)
{
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        double s = 0.0;
        for (int i = 0; i < nblocks; i++) s += partial[i];
        fval[0] += 0.5 * s;
    }
}

static int fill_args(const b2fwi_grid *g, const b2fwi_res2d_plan *p, Res2dArgs *a)
{
    Layout L;
    int rc = make_layout(g, &L);
    if (rc) return rc;
    B2_CHECK_ARG(g->ndim == 2, "the resident engine is 2-D only");
    B2_CHECK_ARG(g->halo == 0, "the resident engine expects halo == 0");
    B2_CHECK_ARG(L.R >= 2 && L.R <= 4, "resident engine: space_order must be 4, 6 or 8 (got %d)", g->space_order);
    memset(a, 0, sizeof(*a));
    a->nx = g->shape[0]; a->nz = g->shape[1];
    a->nzq = (a->nz + 3) / 4;
    a->C = p->cluster; a->rows_cta = p->rows_cta; a->G = p->groups; a->threads = p->threads;
    a->tile_rows = p->tile_rows;
    a->wx0 = p->wx0; a->wx1 = p->wx1; a->wq0 = p->wq0; a->wq1 = p->wq1;
    a->sr = L.sr;
    StepArgs w;
    memset(&w, 0, sizeof(w));
    fill_stencil_weights(L, &w);
    a->c0 = w.c0; a->c0_lo = w.c0_lo;
    for (int k = 0; k <= 4; k++) { a->cx[k] = w.cr[k]; a->cz[k] = w.cz[k]; }
    // consistency of the plan with the grid
    B2_CHECK_ARG(a->C >= 1 && a->C <= 8, "cluster size %d", a->C);
    B2_CHECK_ARG(a->rows_cta * a->C >= a->nx && a->rows_cta * (a->C - 1) < a->nx, "rows_cta %d does not tile nx %d",
                 a->rows_cta, a->nx);
    B2_CHECK_ARG(a->nx - a->rows_cta * (a->C - 1) >= L.R, "last CTA has fewer than R rows");
    B2_CHECK_ARG(a->G * p->rows_per_thread >= a->rows_cta && a->tile_rows == a->G * p->rows_per_thread + 2 * L.R,
                 "bad group / tile geometry");
    B2_CHECK_ARG(a->threads >= a->nzq * a->G && a->threads % 32 == 0 && a->threads <= max_threads_for(p->rows_per_thread),
                 "bad thread count %d", a->threads);
    B2_CHECK_ARG(a->wx0 >= 0 && a->wx1 <= a->nx && a->wx0 < a->wx1 && a->wq0 >= 0 && a->wq1 <= a->nzq && a->wq0 < a->wq1,
                 "bad window");
    B2_CHECK_ARG((int)res2d_smem_bytes(*a, p->rows_per_thread) <= kMaxSmem, "plan exceeds shared memory");
    return 0;
}

static void fill_maps(const b2fwi_res2d_maps *m, Res2dArgs *a)
{
    a->inj_desc = m->inj_desc; a->inj_cptr = m->inj_cptr; a->inj_pt = m->inj_pt; a->inj_w = m->inj_w;
    a->thr_mask = (const unsigned long long *)m->thr_mask; a->thr_base = m->thr_base;
    a->itp_desc = m->itp_desc; a->itp_pt = m->itp_pt; a->itp_off = m->itp_off; a->itp_w = m->itp_w;
}

}  // namespace b2fwi

using namespace b2fwi;

extern "C" {

int b2fwi_res2d_plan_model(const b2fwi_grid *g, int32_t nbl, int32_t min_cluster, b2fwi_res2d_plan *out)
{
    Layout L;
    int rc = make_layout(g, &L);
    if (rc) return rc;
    B2_CHECK_ARG(out != nullptr, "plan_out is NULL");
    if (g->ndim != 2 || L.R < 2 || L.R > 4 || g->halo != 0) {
        set_error("resident engine: needs a 2-D grid, halo 0 and space_order 4, 6 or 8");
        return B2FWI_EUNSUPPORTED;
    }
    const int nx = g->shape[0], nz = g->shape[1], nzq = (nz + 3) / 4;
    B2_CHECK_ARG(nbl >= 0 && 2 * nbl < nx && 2 * nbl < nz, "bad nbl %d", nbl);
    if (min_cluster < 1) min_cluster = 1;
    const int Ps[3] = {8, 12, 16};
    for (int C = min_cluster; C <= 8; C++) {
        const int rows_cta = (nx + C - 1) / C;
        if (rows_cta * (C - 1) >= nx) continue;
        if (nx - rows_cta * (C - 1) < L.R || rows_cta < 2 * L.R) continue;
        for (int ip = 0; ip < 3; ip++) {
            const int P = Ps[ip];
            const int G = (rows_cta + P - 1) / P;
            const int threads = (nzq * G + 31) / 32 * 32;
            if (threads > max_threads_for(P)) continue;
            Res2dArgs a;
            memset(&a, 0, sizeof(a));
            a.nzq = nzq; a.rows_cta = rows_cta; a.G = G; a.tile_rows = G * P + 2 * L.R;
            a.wq0 = nbl / 4; a.wq1 = (nz - nbl + 3) / 4;
            const size_t smem = res2d_smem_bytes(a, P);
            if (smem > (size_t)kMaxSmem) continue;
            out->cluster = C; out->rows_per_thread = P; out->groups = G; out->threads = threads;
            out->rows_cta = rows_cta; out->tile_rows = a.tile_rows; out->smem_bytes = (int32_t)smem;
            out->wx0 = nbl; out->wx1 = nx - nbl; out->wq0 = a.wq0; out->wq1 = a.wq1;
            return 0;
        }
    }
    set_error("resident engine: grid %dx%d (space_order %d) does not fit a cluster of <= 8 SMs", nx, nz, g->space_order);
    return B2FWI_EUNSUPPORTED;
}

int b2fwi_res2d_max_active_clusters(const b2fwi_grid *g, const b2fwi_res2d_plan *plan, int32_t *out)
{
    B2_CHECK_ARG(plan && out, "NULL argument");
    Res2dArgs a;
    int rc = fill_args(g, plan, &a);
    if (rc) return rc;
    int n = 0;
    rc = res2d_max_clusters(a, g->space_order / 2, plan->rows_per_thread, &n);
    *out = n;
    return rc;
}

int b2fwi_res2d_prepare(const b2fwi_grid *g, const float *vp, float dt, float *B_out, void *stream)
{
    Layout L;
    int rc = make_layout(g, &L);
    if (rc) return rc;
    B2_CHECK_ARG(vp && B_out && dt > 0.f, "bad argument");
    bcoef_kernel<<<(unsigned)((L.elems + 255) / 256), 256, 0, (cudaStream_t)stream>>>(vp, (double)dt, B_out, L.elems);
    B2_CUDA(cudaGetLastError());
    count_launch();
    return 0;
}

int b2fwi_res2d_forward(const b2fwi_grid *g, const b2fwi_res2d_plan *plan, const float *B, const float *sx,
                        const float *sz, float dt, int32_t nt, int32_t time_m, int32_t time_M, int32_t nshots,
                        const float *src, int32_t nsrc, const b2fwi_res2d_maps *maps,
                        float *rec, int32_t nrec, float *hist, float *illum_out, void *stream)
{
    B2_CHECK_ARG(plan && B && sx && sz && maps && src, "NULL argument");
    B2_CHECK_ARG(nt >= 3 && time_m >= 1 && time_M <= nt - 2 && time_m <= time_M, "bad time range [%d, %d], nt=%d", time_m,
                 time_M, nt);
    B2_CHECK_ARG(nshots >= 1, "nshots = %d", nshots);
    Res2dArgs a;
    int rc = fill_args(g, plan, &a);
    if (rc) return rc;
    fill_maps(maps, &a);
    a.nt = nt; a.time_m = time_m; a.time_M = time_M; a.inv_dt2 = 1.f / (dt * dt);
    a.B = B; a.sx = sx; a.sz = sz;
    a.nshots = nshots; a.vals = src; a.nvals = nsrc; a.vals_shot_stride = (int64_t)nt * nsrc;
    a.rec = rec; a.nrec = nrec;
    a.hist = hist; a.hist_t0 = time_m;
    a.hist_t_stride = (int64_t)(a.wx1 - a.wx0) * (a.wq1 - a.wq0) * 4;
    a.hist_shot_stride = a.hist_t_stride * (time_M - time_m + 1);
    a.out = illum_out;
    return launch_res2d(a, g->space_order / 2, plan->rows_per_thread, 0, (cudaStream_t)stream);
}

int b2fwi_res2d_gradient(const b2fwi_grid *g, const b2fwi_res2d_plan *plan, const float *B, const float *sx,
                         const float *sz, float dt, int32_t nt, int32_t time_m, int32_t time_M, int32_t nshots,
                         const float *res, int32_t nrec, const b2fwi_res2d_maps *maps,
                         const float *hist, float *grad_out, void *stream)
{
    B2_CHECK_ARG(plan && B && sx && sz && maps && res && hist && grad_out, "NULL argument");
    B2_CHECK_ARG(nt >= 3 && time_m >= 1 && time_M <= nt - 2 && time_m <= time_M, "bad time range [%d, %d], nt=%d", time_m,
                 time_M, nt);
    B2_CHECK_ARG(nshots >= 1, "nshots = %d", nshots);
    Res2dArgs a;
    int rc = fill_args(g, plan, &a);
    if (rc) return rc;
    fill_maps(maps, &a);
    a.nt = nt; a.time_m = time_m; a.time_M = time_M; a.inv_dt2 = 1.f / (dt * dt);
    a.B = B; a.sx = sx; a.sz = sz;
    a.nshots = nshots; a.vals = res; a.nvals = nrec; a.vals_shot_stride = (int64_t)nt * nrec;
    a.rec = nullptr; a.nrec = nrec;
    a.hist = const_cast<float *>(hist); a.hist_t0 = time_m;
    a.hist_t_stride = (int64_t)(a.wx1 - a.wx0) * (a.wq1 - a.wq0) * 4;
    a.hist_shot_stride = a.hist_t_stride * (time_M - time_m + 1);
    a.out = grad_out;
    return launch_res2d(a, g->space_order / 2, plan->rows_per_thread, 1, (cudaStream_t)stream);
}

int b2fwi_window_mask_accumulate(int32_t nx, int32_t nz, const float *field, int64_t row_stride, int32_t col0,
                                 const double *mask, double *out, void *stream)
{
    B2_CHECK_ARG(nx > 0 && nz > 0 && field && out, "bad argument");
    window_mask_acc_kernel<<<(nx * nz + 127) / 128, 128, 0, (cudaStream_t)stream>>>(nx, nz, field, row_stride, col0, mask,
                                                                                  out);
    B2_CUDA(cudaGetLastError());
    count_launch();
    return 0;
}

int b2fwi_window_mask_accumulate_batch(int32_t nshots, int32_t nx, int32_t nz, const float *field, int64_t shot_stride,
                                       int64_t row_stride, int32_t col0, const double *mask, double *out, void *stream)
{
    B2_CHECK_ARG(nshots > 0 && nx > 0 && nz > 0 && field && out, "bad argument");
    window_mask_acc_batch_kernel<<<(nx * nz + 127) / 128, 128, 0, (cudaStream_t)stream>>>(
        nshots, nx, nz, field, shot_stride, row_stride, col0, mask, out);
    B2_CUDA(cudaGetLastError());
    count_launch();
    return 0;
}

int b2fwi_l2_misfit(const float *syn, const float *obs, const float *dw, int64_t n, float *residual_out,
                    double *fval_out, double *scratch, void *stream)
{
    B2_CHECK_ARG(syn && obs && residual_out && fval_out && scratch && n > 0, "bad argument");
    int nblocks = (int)((n + 255) / 256);
    if (nblocks > 1024) nblocks = 1024;
    l2_stage1<<<nblocks, 256, 0, (cudaStream_t)stream>>>(syn, obs, dw, n, residual_out, scratch);
    l2_stage2<<<1, 32, 0, (cudaStream_t)stream>>>(scratch, nblocks, fval_out);
    B2_CUDA(cudaGetLastError());
    count_launch(2);
    return 0;
}

}  // extern "C"

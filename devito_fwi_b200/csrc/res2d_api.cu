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

static int max_threads_for(int P) { return P <= RES2D_LAT_MAXP ? 512 : P <= 8 ? 480 : P <= 12 ? 384 : 288; }
static const int kMaxCluster = 16;     // 9..16 = non-portable cluster sizes (cudaFuncAttributeNonPortableClusterSizeAllowed)

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


// ------------------------------------------------------------------------------------------------
// 1-D quadratic-Wasserstein misfit, trace by trace (misfit/misfit.py:20-67, qWasserstein(method='1d',
// trans_type='linear')): signals are shifted positive by c = gamma * max(0, -min) (one c per shot record),
// normalised to unit mass, T = G^-1(F) by linear interpolation of the cumulative distributions,
//   loss = 0.5 sum (t - T)^2 mu,   grad = (cumsum(t - T) - sum(t - T) - <.., mu>) / mass.
// Follows the reference's numeric types: fp32 signals and fp32 running sums (np.cumsum), fp64 from np.interp on.
__global__ void w1d_min_kernel(const float *__restrict__ syn, const float *__restrict__ obs, const float *__restrict__ dw,
                               int64_t n_per_shot, float *__restrict__ cmin)
{
    __shared__ float sh[256];
    const int shot = blockIdx.x;
    const int64_t base = (int64_t)shot * n_per_shot;
    float m = 3.4e38f;
    for (int64_t i = threadIdx.x; i < n_per_shot; i += blockDim.x) {
        const float d = dw ? dw[base + i] : 0.f;
        m = fminf(m, fminf(syn[base + i] - d, obs[base + i] - d));
    }
    sh[threadIdx.x] = m;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if (threadIdx.x < s) sh[threadIdx.x] = fminf(sh[threadIdx.x], sh[threadIdx.x + s]);
        __syncthreads();
    }
    if (threadIdx.x == 0) cmin[shot] = sh[0];
}

// numpy's float32 `sum()` (pairwise summation, 8 accumulators per <=128-element block, numpy/core/src/umath/
// loops_utils.h) reproduced exactly: the trace mass enters the normalisation, and a 1-ulp difference in it
// shows up as 3e-5 in the adjoint source through the inverse-CDF interpolation.
struct ShiftedTrace {
    const float *a, *dw;
    int64_t stride;
    float c;
    __device__ float operator[](int k) const
    {
        const float d = dw ? dw[(int64_t)k * stride] : 0.f;
        return __fadd_rn(__fsub_rn(a[(int64_t)k * stride], d), c);
    }
};

static __device__ float pairwise_sum_f32(const ShiftedTrace &v, int i0, int n)
{
    if (n < 8) {
        float res = 0.f;
        for (int i = 0; i < n; i++) res = __fadd_rn(res, v[i0 + i]);
        return res;
    }
    if (n <= 128) {
        float r[8];
        for (int j = 0; j < 8; j++) r[j] = v[i0 + j];
        int i;
        for (i = 8; i < n - (n % 8); i += 8)
            for (int j = 0; j < 8; j++) r[j] = __fadd_rn(r[j], v[i0 + i + j]);
        float res = __fadd_rn(__fadd_rn(__fadd_rn(r[0], r[1]), __fadd_rn(r[2], r[3])),
                              __fadd_rn(__fadd_rn(r[4], r[5]), __fadd_rn(r[6], r[7])));
        for (; i < n; i++) res = __fadd_rn(res, v[i0 + i]);
        return res;
    }
    int n2 = n / 2;
    n2 -= n2 % 8;
    return __fadd_rn(pairwise_sum_f32(v, i0, n2), pairwise_sum_f32(v, i0 + n2, n - n2));
}

__global__ void w1d_trace_kernel(const float *__restrict__ syn, const float *__restrict__ obs, const float *__restrict__ dw,
                                 int nt, int nrec, int nshots, double gamma, const float *__restrict__ cmin,
                                 float *__restrict__ G, double *__restrict__ D, float *__restrict__ grad_out,
                                 double *__restrict__ loss)
{
    const int tr = blockIdx.x * blockDim.x + threadIdx.x;       // global trace index: shot * nrec + receiver
    if (tr >= nshots * nrec) return;
    const int shot = tr / nrec, ir = tr - shot * nrec;
    const int64_t base = (int64_t)shot * nt * nrec + ir;        // element (k, ir) at base + k*nrec
    const float mn = cmin[shot];
    // c = (-min) * gamma evaluated in fp32, as numpy does for an fp32 scalar times a python float (NEP 50)
    const float c = (mn < 0.f) ? __fmul_rn(-mn, (float)gamma) : 0.f;
    float *g = G + (int64_t)tr;                                  // scratch, element k at g[k * ntraces]
    double *dd = D + (int64_t)tr;
    const int64_t ntr = (int64_t)nshots * nrec;
    const ShiftedTrace vf{syn + base, dw ? dw + base : nullptr, nrec, c}, vg{obs + base, dw ? dw + base : nullptr, nrec, c};
    const float mass_f = pairwise_sum_f32(vf, 0, nt), mass_g = pairwise_sum_f32(vg, 0, nt);
    // G = cumsum(nu) in fp32
    float acc = 0.f;
    for (int k = 0; k < nt; k++) {
        const float d = dw ? dw[base + (int64_t)k * nrec] : 0.f;
        acc += ((obs[base + (int64_t)k * nrec] - d) + c) / mass_g;
        g[(int64_t)k * ntr] = acc;
    }
    // sweep F = cumsum(mu); T = interp(F, G, t); both sequences are non-decreasing -> one forward pointer
    const double h = 1.0 / (double)(nt - 1);
    float F = 0.f;
    int j = 0;                                                   // largest j with G[j] <= F (searchsorted side='right' - 1)
    const float G0 = g[0], Glast = g[(int64_t)(nt - 1) * ntr];
    double lsum = 0.0, dsum = 0.0;
    for (int k = 0; k < nt; k++) {
        const float d = dw ? dw[base + (int64_t)k * nrec] : 0.f;
        const float mu = ((syn[base + (int64_t)k * nrec] - d) + c) / mass_f;
        F += mu;
        double T;
        if (F <= G0) T = 0.0;                 // np.interp: left value below the first knot
        else if (F >= Glast) T = 1.0;         // right value (t[-1]) at / above the last knot
        else {
            while (j + 1 < nt - 1 && g[(int64_t)(j + 1) * ntr] <= F) j++;
            const double x0 = (double)g[(int64_t)j * ntr], x1 = (double)g[(int64_t)(j + 1) * ntr];
            const double t0 = j * h;
            T = (x1 > x0) ? ((double)F - x0) * (h / (x1 - x0)) + t0 : t0;
        }
        const double dk = k * h - T;
        dd[(int64_t)k * ntr] = dk;
        lsum += dk * dk * (double)mu;
        dsum += dk;
    }
    loss[tr] = 0.5 * lsum;
    // grad = cumsum(d) - sum(d);  grad = (grad - sum(grad * mu)) / mass
    double run = 0.0, gm = 0.0;
    for (int k = 0; k < nt; k++) {
        const float d = dw ? dw[base + (int64_t)k * nrec] : 0.f;
        const float mu = ((syn[base + (int64_t)k * nrec] - d) + c) / mass_f;
        run += dd[(int64_t)k * ntr];
        gm += (run - dsum) * (double)mu;
    }
    run = 0.0;
    for (int k = 0; k < nt; k++) {
        run += dd[(int64_t)k * ntr];
        grad_out[base + (int64_t)k * nrec] = (float)(((run - dsum) - gm) / (double)mass_f);
    }
}

__global__ void sum_doubles_kernel(const double *__restrict__ v, int n, double *__restrict__ out)
{
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        double s = 0.0;
        for (int i = 0; i < n; i++) s += v[i];
        out[0] += s;
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
    B2_CHECK_ARG(a->C >= 1 && a->C <= kMaxCluster, "cluster size %d", a->C);
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
    B2_CHECK_ARG(p->tile_pitch == (p->rows_per_thread <= RES2D_LAT_MAXP ? 4 * res2d_lat_pitch_quads(a->nzq) : 4 * (a->nzq + 2)),
                 "tile pitch %d does not match the kernel's", p->tile_pitch);
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

// plan for exactly (C CTAs per shot, P rows per thread); B2FWI_EUNSUPPORTED when it does not fit
static int plan_exact(const b2fwi_grid *g, const Layout &L, int nbl, int C, int P, b2fwi_res2d_plan *out)
{
    const int nx = g->shape[0], nz = g->shape[1], nzq = (nz + 3) / 4;
    const bool lat = P <= RES2D_LAT_MAXP;
    if (C < 1 || C > kMaxCluster || !(P == 3 || P == 4 || P == 8 || P == 12 || P == 16)) return B2FWI_EUNSUPPORTED;
    const int rows_cta = (nx + C - 1) / C;
    if (rows_cta * (C - 1) >= nx) return B2FWI_EUNSUPPORTED;
    if (nx - rows_cta * (C - 1) < L.R || rows_cta < 2 * L.R) return B2FWI_EUNSUPPORTED;
    if (lat && res2d_lat_pitch_quads(nzq) == 0) return B2FWI_EUNSUPPORTED;
    const int G = (rows_cta + P - 1) / P;
    const int threads = (nzq * G + 31) / 32 * 32;
    if (threads > max_threads_for(P)) return B2FWI_EUNSUPPORTED;
    Res2dArgs a;
    memset(&a, 0, sizeof(a));
    a.nzq = nzq; a.rows_cta = rows_cta; a.G = G; a.tile_rows = G * P + 2 * L.R;
    a.wq0 = nbl / 4; a.wq1 = (nz - nbl + 3) / 4;
    const size_t smem = res2d_smem_bytes(a, P);
    if (smem > (size_t)kMaxSmem) return B2FWI_EUNSUPPORTED;
    out->cluster = C; out->rows_per_thread = P; out->groups = G; out->threads = threads;
    out->rows_cta = rows_cta; out->tile_rows = a.tile_rows; out->smem_bytes = (int32_t)smem;
    out->wx0 = nbl; out->wx1 = nx - nbl; out->wq0 = a.wq0; out->wq1 = a.wq1;
    out->tile_pitch = lat ? 4 * res2d_lat_pitch_quads(nzq) : 4 * (nzq + 2);
    return 0;
}

static int plan_prologue(const b2fwi_grid *g, int32_t nbl, b2fwi_res2d_plan *out, Layout *L)
{
    int rc = make_layout(g, L);
    if (rc) return rc;
    B2_CHECK_ARG(out != nullptr, "plan_out is NULL");
    if (g->ndim != 2 || L->R < 2 || L->R > 4 || g->halo != 0) {
        set_error("resident engine: needs a 2-D grid, halo 0 and space_order 4, 6 or 8");
        return B2FWI_EUNSUPPORTED;
    }
    B2_CHECK_ARG(nbl >= 0 && 2 * nbl < g->shape[0] && 2 * nbl < g->shape[1], "bad nbl %d", nbl);
    return 0;
}

int b2fwi_res2d_plan_model(const b2fwi_grid *g, int32_t nbl, int32_t min_cluster, int32_t min_rows_per_thread,
                           b2fwi_res2d_plan *out)
{
    Layout L;
    int rc = plan_prologue(g, nbl, out, &L);
    if (rc) return rc;
    if (min_cluster < 1) min_cluster = 1;
    // rows per thread: the shortest strip whose thread count fits (time per step goes with warps per scheduler x
    // rows per thread; longer strips only amortise the 2R-row window fill)
    const int Ps[4] = {4, 8, 12, 16};
    for (int C = min_cluster; C <= kMaxCluster; C++)
        for (int ip = 0; ip < 4; ip++)
            if (Ps[ip] >= min_rows_per_thread && plan_exact(g, L, nbl, C, Ps[ip], out) == 0) return 0;
    set_error("resident engine: grid %dx%d (space_order %d) does not fit a cluster of <= 16 SMs", g->shape[0], g->shape[1],
              g->space_order);
    return B2FWI_EUNSUPPORTED;
}

int b2fwi_res2d_plan_exact(const b2fwi_grid *g, int32_t nbl, int32_t cluster, int32_t rows_per_thread,
                           b2fwi_res2d_plan *out)
{
    Layout L;
    int rc = plan_prologue(g, nbl, out, &L);
    if (rc) return rc;
    rc = plan_exact(g, L, nbl, cluster, rows_per_thread, out);
    if (rc) set_error("resident engine: no plan with %d CTAs per shot and %d rows per thread", cluster, rows_per_thread);
    return rc;
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

int b2fwi_w1d_misfit(const float *syn, const float *obs, const float *dw, int32_t nt, int32_t nrec, int32_t nshots,
                     double gamma, float *adjsrc_out, double *fval_out, void *scratch, void *stream)
{
    B2_CHECK_ARG(syn && obs && adjsrc_out && fval_out && scratch && nt >= 2 && nrec >= 1 && nshots >= 1, "bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t ntr = (int64_t)nshots * nrec;
    // scratch layout: D[nt*ntr] doubles | loss[ntr] doubles | G[nt*ntr] floats | cmin[nshots] floats
    double *D = reinterpret_cast<double *>(scratch);
    double *loss = D + (int64_t)nt * ntr;
    float *G = reinterpret_cast<float *>(loss + ntr);
    float *cmin = G + (int64_t)nt * ntr;
    w1d_min_kernel<<<nshots, 256, 0, st>>>(syn, obs, dw, (int64_t)nt * nrec, cmin);
    w1d_trace_kernel<<<(unsigned)((ntr + 63) / 64), 64, 0, st>>>(syn, obs, dw, nt, nrec, nshots, gamma, cmin, G, D,
                                                                 adjsrc_out, loss);
    sum_doubles_kernel<<<1, 32, 0, st>>>(loss, (int)ntr, fval_out);
    B2_CUDA(cudaGetLastError());
    count_launch(3);
    return 0;
}

int64_t b2fwi_w1d_scratch_bytes(int32_t nt, int32_t nrec, int32_t nshots)
{
    const int64_t ntr = (int64_t)nshots * nrec;
    return ((int64_t)nt * ntr + ntr) * 8 + ((int64_t)nt * ntr + nshots) * 4 + 64;
}

}  // extern "C"

/*
 * oracle/qw2d/dct_shim.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * (1) The two FFTW r2r kinds of fftw3.h (see there), evaluated directly from their definitions: separable,
 *     O(n^2) per line with cosine tables, accumulated in double and rounded once to float (at least as accurate
 *     as FFTW's single-precision FFT; the reference is only defined up to that rounding).
 * (2) qw2d_ref_gradient(): the call sequence of the reference's driver program misfit/QW2D/src/w2.c:36-58
 *     (alloc_fotSpace_2d, init_fotSpace_2d(syn, obs), fotGradient2d on adj = 1) as a library entry point, so that
 *     tests can run the reference's own fot2d.c in-process instead of through files and a subprocess
 *     (misfit/bfm.py:145-193).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "fot2d.h"

struct oracle_dct_plan {
    int n0, n1;
    float *in, *out;
    fftwf_r2r_kind k0, k1;
    double *c0, *c1;   /* cos tables [n][n]: c[k*n + j] */
    double *tmp;
};

static double *cos_table(int n, fftwf_r2r_kind kind)
{
    double *c = (double *)malloc(sizeof(double) * (size_t)n * n);
    for (int k = 0; k < n; k++)
        for (int j = 0; j < n; j++) {
            if (kind == FFTW_REDFT10) c[(size_t)k * n + j] = 2.0 * cos(M_PI * (j + 0.5) * k / n);
            else c[(size_t)k * n + j] = (j == 0) ? 1.0 : 2.0 * cos(M_PI * j * (k + 0.5) / n);
        }
    return c;
}

fftwf_plan fftwf_plan_r2r_2d(int n0, int n1, float *in, float *out, fftwf_r2r_kind kind0, fftwf_r2r_kind kind1,
                             unsigned flags)
{
    (void)flags;
    struct oracle_dct_plan *p = (struct oracle_dct_plan *)malloc(sizeof(*p));
    p->n0 = n0; p->n1 = n1; p->in = in; p->out = out; p->k0 = kind0; p->k1 = kind1;
    p->c0 = cos_table(n0, kind0);
    p->c1 = cos_table(n1, kind1);
    p->tmp = (double *)malloc(sizeof(double) * (size_t)n0 * n1);
    return p;
}

void fftwf_execute(const fftwf_plan p)
{
    const int n0 = p->n0, n1 = p->n1;
    /* along the contiguous dimension (length n1) */
#pragma omp parallel for schedule(static)
    for (int i = 0; i < n0; i++)
        for (int k = 0; k < n1; k++) {
            const double *c = p->c1 + (size_t)k * n1;
            const float *x = p->in + (size_t)i * n1;
            double s = 0.0;
            for (int j = 0; j < n1; j++) s += c[j] * (double)x[j];
            p->tmp[(size_t)i * n1 + k] = s;
        }
    /* along the slow dimension (length n0) */
#pragma omp parallel for schedule(static)
    for (int k = 0; k < n0; k++) {
        const double *c = p->c0 + (size_t)k * n0;
        for (int j = 0; j < n1; j++) {
            double s = 0.0;
            for (int i = 0; i < n0; i++) s += c[i] * p->tmp[(size_t)i * n1 + j];
            p->out[(size_t)k * n1 + j] = (float)s;
        }
    }
}

void fftwf_destroy_plan(fftwf_plan p)
{
    free(p->c0); free(p->c1); free(p->tmp); free(p);
}

/* misfit/QW2D/src/w2.c:8-75 without the file I/O. n1 = fastest dimension (python: f.shape[1]). Returns the loss. */
float qw2d_ref_gradient(int n1, int n2, int niter, float step_scale, const float *syn, const float *obs, float *adj)
{
    struct fotSpace otspace;
    otspace.nIter = niter;
    otspace.step_scale = step_scale;
    alloc_fotSpace_2d(&otspace, n1, n2);
    for (int i = 0; i < n1 * n2; i++) adj[i] = 1.0f;
    init_fotSpace_2d(&otspace, n1, n2, (float *)syn, (float *)obs);
    float w = fotGradient2d(&otspace, adj, n1, n2, 0);
    destroy_fotSpace_2d(&otspace);
    return w;
}

/* ---- single steps of the solver, for stage-by-stage comparisons (tests/test_gpu_qw2d.py) ---- */
void qw2d_ref_dual(int n1, int n2, const float *u, float *dual)              /* compute_2d_dual, fot2d.c:157-183 */
{
    struct convex_hull hull;
    alloc_hull(&hull, n1 > n2 ? n1 : n2);
    compute_2d_dual(dual, (float *)u, &hull, n1, n2);
    destroy_hull(&hull);
}

float qw2d_ref_update(int n1, int n2, float *pot, const float *rho, const float *other, float sigma)   /* fot2d.c:479-503 */
{
    struct poisson_solver ps = create_poisson_solver_workspace2d(n1, n2);
    float h1 = update_potential(ps, pot, (float *)rho, (float *)other, sigma, n1 * n2);
    destroy_poisson_solver(ps);
    return h1;
}

void qw2d_ref_push(int n1, int n2, const float *pot, const float *dens, float *rho)   /* fot2d.c:290-322,398-478 */
{
    float *xm = (float *)calloc((size_t)(n1 + 1) * (n2 + 1), sizeof(float));
    float *ym = (float *)calloc((size_t)(n1 + 1) * (n2 + 1), sizeof(float));
    calc_pushforward_map(xm, ym, (float *)pot, n1, n2);
    sampling_pushforward(rho, (float *)dens, xm, ym, n1, n2, 1.0f);
    free(xm); free(ym);
}

float qw2d_ref_w2(int n1, int n2, const float *phi, const float *dual, const float *mu, const float *nu)   /* fot2d.c:519-531 */
{
    return compute_w2((float *)phi, (float *)dual, (float *)mu, (float *)nu, n1, n2);
}

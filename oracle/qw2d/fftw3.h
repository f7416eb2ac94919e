/*
 * oracle/qw2d/fftw3.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Stand-in for the one third-party dependency of the reference's QW2D solver (misfit/QW2D/src/fot2d.h:14 includes
 * <fftw3.h>; misfit/QW2D/src/Makefile links -lfftw3f, no version pinned; FFTW 3.3.x at the time of the reference).
 * libfftw3f is not installed in this image, so the recipe in oracle/qw2d/Makefile compiles the reference's own,
 * unmodified sources (fot2d.c, normalize.c) against this header and dct_shim.c, which restates the two published FFTW
 * r2r kinds the solver uses (FFTW manual, "1d Real-even DFTs (DCTs)"):
 *   FFTW_REDFT10 (DCT-II)   Y_k = 2 sum_{j=0}^{n-1} X_j cos(pi (j + 1/2) k / n)
 *   FFTW_REDFT01 (DCT-III)  Y_k = X_0 + 2 sum_{j=1}^{n-1} X_j cos(pi j (k + 1/2) / n)
 * applied separably along both dimensions (fftwf_plan_r2r_2d, row-major n0 x n1), unnormalised.
 * Only what fot2d.c calls is declared (fot2d.c:27-32,44-45,487,494).
 */
#ifndef ORACLE_QW2D_FFTW3_SHIM_H
#define ORACLE_QW2D_FFTW3_SHIM_H

typedef enum { FFTW_REDFT10 = 5, FFTW_REDFT01 = 4 } fftwf_r2r_kind;
#define FFTW_MEASURE (0U)

typedef struct oracle_dct_plan *fftwf_plan;

fftwf_plan fftwf_plan_r2r_2d(int n0, int n1, float *in, float *out, fftwf_r2r_kind kind0, fftwf_r2r_kind kind1,
                             unsigned flags);
void fftwf_execute(const fftwf_plan p);
void fftwf_destroy_plan(fftwf_plan p);

#endif

/*
 * oracle/fwi_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement (plain C + OpenMP, fp32 and fp64) of the FWI-gradient hot path of
 * LongyanU/devito-fwi: 2-D/3-D isotropic acoustic OT2 time stepping with sponge damping,
 * multilinear source injection / receiver interpolation, adjoint sweep and zero-lag
 * imaging condition. Only tests/, __graft_entry__.smoke() and the cpu_baseline /
 * --impl reference legs of bench.py may load this library; the product path
 * (devito_fwi_b200/) never does.
 *
 * Why a restatement: the reference executes this path through the third-party package
 * `devito` (README.md:6 -- un-pinned git master, vintage ~v4.3), which is neither vendored in
 * /root/reference nor installable offline, so there is no reference binary to build
 * (oracle/_ref does not exist for this repo). The algorithm restated here is the one
 * the in-tree symbolic definitions denote (file:line citations in fwi_oracle_body.inc)
 * and SURVEY.md Appendix A spells out. It is PINNED by tests/test_oracle_kat.py against
 * every known-answer value the reference holds for the path:
 *   seismic/inversion/fwi.py:95-97,121      objective 39113, grad min/max -821/2442, 5th GD objective 3828 (+-10)
 *   seismic/acoustic/acoustic_example.py:75-79  3-D |rec|_2 = 459.1678 (rtol 1e-3, fp64)
 *   seismic/acoustic/accuracy.ipynb cells 14,16 trace min/max -5.349877e-03/+8.529867e-03, RMS err 1.1265e-05
 *   seismic/acoustic/acoustic_example.py:75-79  free surface, fp32: |rec|_2 = 369.955 (rtol 1e-3)
 * The named Marmousi / circle(so=6) / 3-D 512^3 shapes have no stored outputs in the
 * reference ("parity unpinned" at those shapes).
 *
 * Build: see oracle/Makefile (gcc -O3 -fopenmp; the *_fast variant adds Devito's own
 * -march=native -ffast-math flag set and is used only as the timed CPU baseline).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORACLE_MAX_R 8

/* The *_fast build (timed CPU baseline only) flushes denormals like Devito's generated code does
 * (sa_01_iso_implementation1.ipynb:1222-1224: _MM_SET_DENORMALS_ZERO_MODE / _MM_SET_FLUSH_ZERO_MODE);
 * without it the exponentially small wavefront tail runs in microcode. MXCSR is per thread, so it is set
 * inside the parallel loops. The strict build (the parity checker) keeps IEEE gradual underflow. */
#if defined(ORACLE_FTZ) && (defined(__x86_64__) || defined(__i386__))
#include <xmmintrin.h>
#define ORACLE_SET_FTZ() _mm_setcsr(_mm_getcsr() | 0x8040)
#else
#define ORACLE_SET_FTZ() ((void)0)
#endif

typedef struct {
    int ndim;        /* 2 or 3 */
    int shape[3];    /* padded grid points per dimension, last dimension contiguous */
    int space_order; /* even, 2..16 */
    double spacing[3];
    double origin[3]; /* padded origin (model.py:100), already rounded through the grid dtype */
    int fs;           /* free surface at index 0 of the last dimension (model.py:102-109, operators.py:8-35) */
    int ot4;          /* kernel='OT4': H = laplace + s^2/12 * biharmonic(1/m) (operators.py:38-56) */
} oracle_grid;

/*
 * Central second-derivative weights of order `so` on a unit grid (field.laplace,
 * operators.py:56): c_k = 2 (-1)^(k+1) (R!)^2 / (k^2 (R-k)! (R+k)!), c_0 = -2 sum_k c_k.
 */
void oracle_laplace_coeffs(int so, double *c)
{
    int R = so / 2;
    double c0 = 0.0;
    for (int k = 1; k <= R; k++) {
        /* (R!)^2 / ((R-k)! (R+k)!) = prod_{j=1..k} (R-k+j)/(R+j) */
        double r = 1.0;
        for (int j = 1; j <= k; j++) r *= (double)(R - k + j) / (double)(R + j);
        c[k] = ((k & 1) ? 2.0 : -2.0) * r / ((double)k * (double)k);
        c0 += c[k];
    }
    c[0] = -2.0 * c0;
}

#define REAL float
#define SUFFIX _f32
#define FLOOR floorf
#include "fwi_oracle_body.inc"
#undef REAL
#undef SUFFIX
#undef FLOOR

#define REAL double
#define SUFFIX _f64
#define FLOOR floor
#include "fwi_oracle_body.inc"
#undef REAL
#undef SUFFIX
#undef FLOOR

/*
 * b2fwi.h -- C ABI of the B200-native acoustic FWI-gradient engine (libb2fwi.so).
 *
 * This is the drop-in boundary for the one hot path of LongyanU/devito-fwi: what the
 * reference reaches through `devito.Operator.apply` -> ctypes -> JIT-generated C
 *     int Forward (struct dataobj *damp_vec, const float dt, struct dataobj *m_vec, ..., int time_M, int time_m, struct profiler*)
 *     int Gradient(...), int Adjoint(...)
 * (reference: seismic/acoustic/wavesolver.py:112,149,203 call sites; generated-code signature in
 * seismic/tutorials/08_snapshotting.ipynb:475 and seismic/self_adjoint/sa_01_iso_implementation1.ipynb:1187-1320).
 *
 * Conventions
 *  - Plain C: pointers + sizes only. No allocation inside the library: the caller owns every buffer.
 *  - All `float*` / index arguments are DEVICE pointers unless the name ends in `_host`.
 *  - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream). Calls are asynchronous
 *    with respect to the host; they are ordered on `stream`.
 *  - Every entry point returns 0 on success, a negative B2FWI_E* code otherwise;
 *    b2fwi_last_error() gives the message of the calling thread's last failure.
 *  - Arithmetic is IEEE fp32 (the reference's default grid dtype, seismic/model.py:92).
 *
 * Device layout of a grid field ("pitched slice"): the padded model grid
 * shape[] = physical shape + 2*nbl (seismic/model.py:101), C order, with the contiguous (last)
 * dimension pitched to a multiple of 32 floats so that every row starts 128-byte aligned, and
 * optionally surrounded by `halo` zero cells on every side (halo = 0 is the normal case: the
 * homogeneous-Dirichlet exterior Devito keeps in its halo region is synthesised by predicated loads).
 * Use b2fwi_field_layout(). Pitch-padding cells must be zero on entry and are kept zero by every kernel.
 */
#ifndef B2FWI_H
#define B2FWI_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B2FWI_VERSION 100

#define B2FWI_OK 0
#define B2FWI_EINVAL (-1)   /* bad argument (shape, order, time range, null pointer) */
#define B2FWI_ECUDA (-2)    /* CUDA runtime error; see b2fwi_last_error() */
#define B2FWI_EUNSUPPORTED (-3)

typedef struct b2fwi_grid {
    int32_t ndim;          /* 2: (x, z)   3: (x, y, z); the last dimension is contiguous */
    int32_t shape[3];      /* padded-grid points per dimension (model.grid.shape) */
    int32_t space_order;   /* 2, 4, 6, 8, 10, 12, 14 or 16 (AcousticWaveSolver(space_order=)) */
    int32_t halo;          /* zero cells around the domain in the device layout; multiple of 4, normally 0 */
    float spacing[3];      /* model.spacing */
    float origin[3];       /* padded origin, model.grid.origin (seismic/model.py:100) */
} b2fwi_grid;

/*
 * Sparse points (SparseTimeFunction coordinates) pre-resolved against the grid: multilinear corner
 * offsets/weights for interpolation and a cell-centric CSR for deterministic injection
 * (replaces the generated inject/interpolate sections, sa_01_iso_implementation1.ipynb:1257-1315).
 * Built on the host by devito_fwi_b200.sparse.SparseMap; all arrays live on the device.
 */
typedef struct b2fwi_sparse {
    int32_t npoint;
    int32_t ncorner;            /* 4 (2-D) or 8 (3-D) */
    const int64_t *corner_off;  /* [npoint*ncorner] element offset into a haloed slice, -1 = outside the grid */
    const float *corner_w;      /* [npoint*ncorner] */
    int32_t ncell;              /* distinct grid cells touched by any point */
    const int64_t *cell_off;    /* [ncell] element offset into a haloed slice */
    const int32_t *cell_ptr;    /* [ncell+1] CSR row pointers into contrib_* */
    const int32_t *contrib_pt;  /* [ncontrib] point index, ascending within a cell */
    const float *contrib_w;     /* [ncontrib] multilinear weight */
} b2fwi_sparse;

/* How the forward wavefield is supplied to the imaging condition of b2fwi_gradient(). */
#define B2FWI_HIST_U 1    /* u_hist[nt][slice]: the saved wavefield itself (TimeFunction(save=nt)) */
#define B2FWI_HIST_D2U 2  /* d2u[nt][slice]: slices of u.dt2 written by b2fwi_forward(d2u_out=...) */

int32_t b2fwi_version(void);
const char *b2fwi_last_error(void);

/* Layout of one haloed slice: element strides per dimension, offset of domain cell (0,..,0), total floats. */
int b2fwi_field_layout(const b2fwi_grid *g, int64_t stride_out[3], int64_t *base_out, int64_t *elems_out);

/*
 * Per-point update coefficients of the OT2 scheme (replaces the in-kernel `m`, `damp` algebra of
 * seismic/acoustic/operators.py:87 / acoustic_time_update_nb.ipynb cell 3):
 *   coef[0] = m / (m + dt*damp),  coef[1] = dt^2 / (m + dt*damp),  m = 1/(vp*vp)   (evaluated in fp64, rounded once)
 * so that  u+ = u + coef0*(u - u-) + coef1*L(u)  ==  [dt^2 L(u) + dt damp u + m(2u - u-)] / (m + dt damp).
 * vp, damp: haloed slices. coef: two consecutive haloed slices.
 */
int b2fwi_prepare_coeffs(const b2fwi_grid *g, const float *vp, const float *damp, float dt, float *coef,
                         void *stream);

/*
 * Forward operator (seismic/acoustic/operators.py:98-140; AcousticWaveSolver.forward, wavesolver.py:76-114).
 *   for time = time_m .. time_M:  u[time+1] = step(u[time], u[time-1]);  u[time+1] += inject(src[time]);
 *                                 rec[time] = interpolate(u[time])
 * u:     save == 0 -> 3 haloed slices, slot = time % 3;  save != 0 -> nt haloed slices (slot = time).
 *        Initial state is read from the caller's u (slots time_m-1, time_m); results are left in place.
 * src:   [nt][src_map->npoint] (may be NULL with npoint == 0);  rec: [nt][rec_map->npoint], rows time_m..time_M written.
 * illum: optional haloed slice, incremented by sum_t u[t]^2 over t = time_m .. time_M+1  (fwi.py:170).
 * d2u_out: optional haloed slices; slice (t - d2u_t0), t = time_m..time_M, receives (u[t-1] - 2u[t] + u[t+1]) / dt^2.
 */
int b2fwi_forward(const b2fwi_grid *g, const float *vp, const float *coef, float dt,
                  int32_t nt, int32_t time_m, int32_t time_M,
                  const float *src, const b2fwi_sparse *src_map,
                  float *rec, const b2fwi_sparse *rec_map,
                  float *u, int32_t save, float *illum, float *d2u_out, int32_t d2u_t0, void *stream);

/*
 * Gradient operator (operators.py:183-225; AcousticWaveSolver.jacobian_adjoint, wavesolver.py:153-205).
 *   for time = time_M .. time_m:  v[time-1] = step(v[time], v[time+1]);  v[time-1] += inject(rec[time]);
 *                                 grad += -u.dt2[time] * v[time]
 * hist/hist_kind: see B2FWI_HIST_*.  hist_t0: time index held by slice 0 of `hist` (0 for a full history;
 *   a checkpoint segment passes its first time index).  v: 3 haloed slices (in/out).  grad: accumulated into.
 */
int b2fwi_gradient(const b2fwi_grid *g, const float *vp, const float *coef, float dt,
                   int32_t nt, int32_t time_m, int32_t time_M,
                   const float *rec, const b2fwi_sparse *rec_map,
                   const float *hist, int32_t hist_kind, int32_t hist_t0,
                   float *v, float *grad, void *stream);

/*
 * Adjoint operator (operators.py:143-180; AcousticWaveSolver.adjoint, wavesolver.py:116-151):
 * as b2fwi_gradient without imaging; srca[time] = interpolate(v[time]) at the source positions.
 */
int b2fwi_adjoint(const b2fwi_grid *g, const float *vp, const float *coef, float dt,
                  int32_t nt, int32_t time_m, int32_t time_M,
                  const float *rec, const b2fwi_sparse *rec_map,
                  float *srca, const b2fwi_sparse *src_map,
                  float *v, void *stream);

/*
 * Per-shot host post-processing of fwi.py:104-129,166-171 moved on device (2-D models):
 *   mask[i,j] = prod_k (1 - exp(-.5*((z_j - c_k0)^2 + (x_i - c_k1)^2) / sigma^2)),  sigma = dx + dz,
 * over the source and all receivers, axes swapped exactly as fix_source_illumination does
 * (np.meshgrid(z, x), fwi.py:115). The product depends on the acquisition geometry only, so it is
 * evaluated once per shot position (fp64, like the reference's numpy arithmetic) and re-used.
 * pts: [npts][2] fp64 positions, source first.  mask_out: dense [nx][nz] fp64, nx = shape[0]-2*nbl.
 */
int b2fwi_geometry_mask(const b2fwi_grid *g, int32_t nbl, const double *pts, int32_t npts,
                        double *mask_out, void *stream);

/*
 * out[i,j] += field[nbl+i, nbl+j] * mask[i,j]   (fwi.py:167-171 crop + mask, fwi.py:195-199 sum over shots).
 * field: haloed slice; mask may be NULL (plain crop); out: dense [nx][nz] fp64 accumulator.
 */
int b2fwi_crop_mask_accumulate(const b2fwi_grid *g, int32_t nbl, const float *field,
                               const double *mask, double *out, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* B2FWI_H */

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

#define B2FWI_VERSION 101

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
    int32_t fs;            /* free surface at index 0 of the last dimension (Model(fs=True), model.py:102-109): the top
                            * rows take the antisymmetric mirror stencil of operators.py:8-35. Streaming engine only. */
    int32_t kernel;        /* 0 = 'OT2'; 1 = 'OT4' (operators.py:38-56): every sweep adds the double-Laplacian term
                            * dt^2/12 * L(vp^2 L(u)) to the update (two more launches per step), the caller passes
                            * dt = 1.73 x critical_dt (wavesolver.py:41-46). Streaming engine, imaging from the saved
                            * wavefield (B2FWI_HIST_U) only; not with fs, not for b2fwi_born. */
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
    /* Optional (3-D, halo 0): tables that let the TMA sweep kernels do injection and interpolation THEMSELVES
     * (operators.py:131-140 is one operator; without them a step is three launches). Contributions are sorted
     * by cell offset, i.e. by (plane, row, z), then by point. NULL / 0: injection and interpolation run as their
     * own small kernels after each sweep. */
    int32_t row_tile;           /* rows per row tile of the pt_* tables (16 = the TMA tile height), 0 = no tables */
    const int32_t *con_rowptr;  /* [np*nr+1] first contribution whose cell lies in (plane, row) */
    const int64_t *con_off;     /* [ncontrib] cell offset of every contribution (cell_off repeated) */
    int32_t max_row_con;        /* most contributions any (plane, row) receives (the kernels stage <= 8, or <= 4 for so > 8) */
    const int32_t *pt_order;    /* [npoint] points sorted by the offset of their first in-grid corner ("home") */
    const int64_t *pt_home;     /* [npoint] that offset, ascending */
    const int32_t *pt_rowptr;   /* [np*nrt+1] first entry of pt_order whose home lies in row tile (plane, rt); nrt = ceil(nr/row_tile) */
    int32_t z_min, z_max;       /* index range of the cells (and homes) per dimension: tiles outside the box have no */
    int32_t r_min, r_max;       /*   sparse work and pay nothing for the fusion (a point source touches 2 x 2 x 2 cells) */
    int32_t p_min, p_max;
} b2fwi_sparse;

/* How the forward wavefield is supplied to the imaging condition of b2fwi_gradient(). */
#define B2FWI_COEF_TAIL 8

#define B2FWI_HIST_U 1    /* u_hist[nt][slice]: the saved wavefield itself (TimeFunction(save=nt)) */
#define B2FWI_HIST_D2U 2  /* d2u[nt][slice]: slices of u.dt2 written by b2fwi_forward(d2u_out=...) */
/* u_hist[nt][slice] again, but one history value per point and step: the imaging sum is taken by parts,
 *   sum_t u.dt2[t] v[t] = sum_t u[t] v.dt2[t] + (u[M+1] v[M] - u[M] v[M+1] - u[m] v[m-1] + u[m-1] v[m]) / dt^2,
 * and v.dt2[t] = (v[t+1] - 2 v[t] + v[t-1]) / dt^2 is formed from the three adjoint levels the sweep holds anyway.
 * The caller owns the boundary term (zero when v starts from rest and u[m-1] = u[m] = 0: the FWI gradient).
 * Checkpoint segments use it: the recompute sweep then writes nothing but the wavefield itself (checkpoint.py). */
#define B2FWI_HIST_UVDT2 3

int32_t b2fwi_version(void);
const char *b2fwi_last_error(void);
/* Number of CUDA kernels this library has launched in the calling process (bench.py's gpu_launches). */
int64_t b2fwi_launch_count(void);
/* Engine switches for A/B runs and the parity tests (process-wide; not thread-safe against running sweeps).
 *   "tma": bit 0 = TMA-staged forward sweep, bit 1 = TMA-staged adjoint + imaging sweep (default 3; 0 = the
 *          register-staged kernels everywhere).
 *   "fuse": sparse operators inside the TMA sweep kernels (service warps) instead of separate launches after each
 *          sweep; bit 0 = injection of small maps (sources), bit 1 = interpolation, bit 2 = injection of any map
 *          (default 7; 0 = three launches per step). Needs the optional tables of b2fwi_sparse.
 *   Returns the previous value (>= 0) or B2FWI_EINVAL. */
int b2fwi_set_option(const char *name, int32_t value);

/* Layout of one haloed slice: element strides per dimension, offset of domain cell (0,..,0), total floats. */
int b2fwi_field_layout(const b2fwi_grid *g, int64_t stride_out[3], int64_t *base_out, int64_t *elems_out);

/*
 * Per-point update coefficients of the OT2 scheme (replaces the in-kernel `m`, `damp` algebra of
 * seismic/acoustic/operators.py:87 / acoustic_time_update_nb.ipynb cell 3):
 *   coef[0] = m / (m + dt*damp),  coef[1] = dt^2 / (m + dt*damp),  m = 1/(vp*vp)   (evaluated in fp64, rounded once)
 * so that  u+ = u + coef0*(u - u-) + coef1*L(u)  ==  [dt^2 L(u) + dt damp u + m(2u - u-)] / (m + dt damp).
 * vp, damp: pitched slices. coef: two consecutive pitched slices followed by B2FWI_COEF_TAIL floats
 * (2*elems + B2FWI_COEF_TAIL in total): the tail receives the index box of the undamped interior, where
 * coef[0] == 1 exactly, so that the sweeps do not read coef[0] there (13 % less HBM traffic on 592^3).
 */
int b2fwi_prepare_coeffs(const b2fwi_grid *g, const float *vp, const float *damp, float dt, float *coef,
                         void *stream);

/*
 * Forward operator (seismic/acoustic/operators.py:98-140; AcousticWaveSolver.forward, wavesolver.py:76-114).
 *   for time = time_m .. time_M:  u[time+1] = step(u[time], u[time-1]);  u[time+1] += inject(src[time]);
 *                                 rec[time] = interpolate(u[time])
 * u:     save == 0 -> 3 haloed slices, slot = time % 3;  save != 0 -> nt haloed slices (slot = time).
 *        Initial state is read from the caller's u (slots time_m-1, time_m); results are left in place.
 *        With save != 0 only slots time_m-1 .. time_M+1 are touched, so a window of the history may be passed as
 *        (buffer - t0 * elems) when buffer[0] holds time level t0: the checkpoint segments of checkpoint.py write
 *        the wavefield of steps ta..tb straight into an S+2-slice buffer this way, and b2fwi_gradient then images
 *        from it (hist_t0 = ta - 1).
 * src:   [nt][src_map->npoint] (may be NULL with npoint == 0);  rec: [nt][rec_map->npoint], rows time_m..time_M written.
 * illum: optional haloed slice, incremented by sum_t u[t]^2 over t = time_m .. time_M, plus u[nt-1]^2 when
 *   time_M == nt-2 (the call that produces the last slice adds it): consecutive time windows tile the sum of
 *   fwi.py:170 exactly.
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
 * Linearised (Born) forward operator (operators.py:228-273 BornOperator; AcousticWaveSolver.jacobian / .born,
 * wavesolver.py:207-242). Per time step, in the reference's expression order:
 *   u[time+1] = step(u[time], u[time-1]);  u[time+1] += inject(src[time]);
 *   U[time+1] = step(U[time], U[time-1]) - c2 * dm * u.dt2[time],   u.dt2 = (u[time-1] - 2u[time] + u[time+1]) / dt^2
 *   rec[time] = interpolate(U[time])
 * (c2 = dt^2 / (m + dt*damp): the source term q = -dm * u.dt2 of iso_stencil solved for U.forward.)
 * dm: haloed slice, perturbation of the squared slowness.  u, U: 3 haloed slices each (ring, in/out).
 * scratch: one haloed slice of workspace (u.dt2 of the current step).
 */
int b2fwi_born(const b2fwi_grid *g, const float *vp, const float *coef, float dt,
               int32_t nt, int32_t time_m, int32_t time_M,
               const float *src, const b2fwi_sparse *src_map,
               float *rec, const b2fwi_sparse *rec_map,
               const float *dm, float *u, float *U, float *scratch, void *stream);

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


/* ------------------------------------------------------------------------------------------------
 * SM-resident 2-D engine: all time steps of all shots of a rank in ONE launch, one thread-block
 * cluster per shot, wavefields in shared memory / registers (devito_fwi_b200/csrc/resident2d.cu).
 * It computes exactly what b2fwi_forward / b2fwi_gradient compute for a zero initial state, for
 * 2-D grids with space_order <= 8 and a separable damping profile (seismic/model.py:31-49), and is
 * what fwi.py:fm_multi / fwi_obj_multi (the shot loops, fwi.py:67-81,183-199) run on.
 * Only the imaging window [wx0,wx1) x quads [wq0,wq1) (the physical domain the reference crops the
 * gradient to, fwi.py:166-167) is accumulated / stored.
 */
typedef struct b2fwi_res2d_plan {
    int32_t cluster;          /* CTAs (SMs) per shot */
    int32_t rows_per_thread;  /* P */
    int32_t groups;           /* row groups per CTA */
    int32_t threads;          /* threads per CTA */
    int32_t rows_cta;         /* grid rows per CTA */
    int32_t tile_rows;
    int32_t smem_bytes;
    int32_t wx0, wx1;         /* window rows, padded-grid coordinates */
    int32_t wq0, wq1;         /* window columns in units of 4 cells (quads) */
    int32_t tile_pitch;       /* floats per row of the CTA's shared-memory tile (b2fwi_res2d_maps.itp_off is built on it) */
} b2fwi_res2d_plan;

/* Host-built maps (devito_fwi_b200.resident), indexed by shot*cluster + rank; device pointers. */
typedef struct b2fwi_res2d_maps {
    const int32_t *inj_desc;   /* [.][2] ncell, cell_base */
    const int32_t *inj_cptr;   /* CSR pointers per cell slot into inj_pt / inj_w */
    const int32_t *inj_pt;     /* point index of each contribution, ascending within a cell */
    const float *inj_w;        /* multilinear weight of each contribution */
    const uint64_t *thr_mask;  /* [.][threads] bit r*4+j set: lane j of row r of the thread's strip is a cell */
    const int32_t *thr_base;   /* [.][threads] first cell slot of the thread */
    const int32_t *itp_desc;   /* [.][2] count, base  (receivers recorded by this CTA; forward only) */
    const int32_t *itp_pt;     /* receiver index */
    const int32_t *itp_off;    /* [.][4] float offsets into the CTA's shared tile, -1 = outside the grid */
    const float *itp_w;        /* [.][4] */
} b2fwi_res2d_maps;

/* Decomposition of a 2-D grid onto clusters; window = the grid minus `nbl` cells on every side.
 * min_cluster: smallest cluster size to try (1..16; sizes above 8 are non-portable cluster sizes, which sm_100
 * offers). Returns B2FWI_EUNSUPPORTED when nothing fits.
 * Strips of 3 or 4 rows per thread run the few-shots kernels (resident2d_lat.cuh), which stage at most 1024 receivers per CTA
 * and a source / residual row span of at most 1024 points per CTA; min_rows_per_thread = 8 plans around it. */
int b2fwi_res2d_plan_model(const b2fwi_grid *g, int32_t nbl, int32_t min_cluster, int32_t min_rows_per_thread,
                           b2fwi_res2d_plan *plan_out);
/* The plan with exactly `cluster` CTAs per shot and `rows_per_thread` in {3, 4, 8, 12, 16}, or B2FWI_EUNSUPPORTED:
 * the host enumerates these and picks by its cost model (devito_fwi_b200/resident.py). */
int b2fwi_res2d_plan_exact(const b2fwi_grid *g, int32_t nbl, int32_t cluster, int32_t rows_per_thread,
                           b2fwi_res2d_plan *plan_out);

/* Number of clusters (= shots) of this plan that the current device keeps resident at the same time
 * (cudaOccupancyMaxActiveClusters); used to pick the cluster size for a given number of shots. */
int b2fwi_res2d_max_active_clusters(const b2fwi_grid *g, const b2fwi_res2d_plan *plan, int32_t *out);

/* B = dt^2 vp^2 (fp64, rounded once) as a pitched slice. */
int b2fwi_res2d_prepare(const b2fwi_grid *g, const float *vp, float dt, float *B_out, void *stream);

/*
 * Forward sweep of `nshots` shots (zero initial state), time_m..time_M:
 *   src [nshots][nt][nsrc] -> rec [nshots][nt][nrec] (rows time_m..time_M written; may be NULL),
 *   hist (nullable) [nshots][time_M-time_m+1][wx1-wx0][(wq1-wq0)*4]: u.dt2 of the window, slice t-time_m,
 *   illum_out (nullable) [nshots][wx1-wx0][(wq1-wq0)*4]: sum_t u[t]^2 of the window (overwritten).
 * sx [shape[0]], sz [4*ceil(shape[1]/4)]: the two parts of damp/dt.
 */
int b2fwi_res2d_forward(const b2fwi_grid *g, const b2fwi_res2d_plan *plan, const float *B, const float *sx,
                        const float *sz, float dt, int32_t nt, int32_t time_m, int32_t time_M, int32_t nshots,
                        const float *src, int32_t nsrc, const b2fwi_res2d_maps *maps,
                        float *rec, int32_t nrec, float *hist, float *illum_out, void *stream);

/*
 * Backward sweep with imaging condition: res [nshots][nt][nrec] injected, hist as written by
 * b2fwi_res2d_forward, grad_out [nshots][wx1-wx0][(wq1-wq0)*4] = -sum_t u.dt2[t] v[t] (overwritten).
 */
int b2fwi_res2d_gradient(const b2fwi_grid *g, const b2fwi_res2d_plan *plan, const float *B, const float *sx,
                         const float *sz, float dt, int32_t nt, int32_t time_m, int32_t time_M, int32_t nshots,
                         const float *res, int32_t nrec, const b2fwi_res2d_maps *maps,
                         const float *hist, float *grad_out, void *stream);

/*
 * out[i,j] += field[i*row_stride + col0 + j] * mask[i,j], i < nx, j < nz: the crop + mute + shot sum of
 * fwi.py:166-171,195-199 for a window-layout field. mask (fp64, nullable) from b2fwi_geometry_mask.
 */
int b2fwi_window_mask_accumulate(int32_t nx, int32_t nz, const float *field, int64_t row_stride, int32_t col0,
                                 const double *mask, double *out, void *stream);

/* Same, summed over `nshots` window-layout fields (shot stride in floats) with per-shot masks
 * mask[nshots][nx][nz]; shots are added in ascending order. */
int b2fwi_window_mask_accumulate_batch(int32_t nshots, int32_t nx, int32_t nz, const float *field, int64_t shot_stride,
                                       int64_t row_stride, int32_t col0, const double *mask, double *out, void *stream);

/*
 * On-device least-squares misfit (misfit/misfit.py:5-9 with the direct-wave subtraction of
 * fwi.py:146-150): residual = (syn - dw) - (obs - dw) in fp32 (dw nullable), fval_out[0] += 0.5*sum residual^2
 * accumulated in fp64 (deterministic two-stage reduction). n = elements.
 */
int b2fwi_l2_misfit(const float *syn, const float *obs, const float *dw, int64_t n, float *residual_out,
                    double *fval_out, double *scratch /* >= 1024 doubles */, void *stream);

/*
 * On-device 1-D quadratic-Wasserstein misfit, trace by trace: qWasserstein(trans_type='linear', method='1d')
 * of misfit/misfit.py:11-104 (with the direct-wave subtraction of fwi.py:146-150, dw nullable).
 * syn, obs, dw, adjsrc_out: [nshots][nt][nrec] fp32; the positivity shift c = gamma*max(0, -min) is taken per shot
 * record as the reference does; fval_out[0] += sum of the trace losses. scratch: b2fwi_w1d_scratch_bytes() bytes.
 */
int b2fwi_w1d_misfit(const float *syn, const float *obs, const float *dw, int32_t nt, int32_t nrec, int32_t nshots,
                     double gamma, float *adjsrc_out, double *fval_out, void *scratch, void *stream);
int64_t b2fwi_w1d_scratch_bytes(int32_t nt, int32_t nrec, int32_t nshots);

/*
 * On-device 2-D quadratic-Wasserstein misfit of whole shot records: qWasserstein(trans_type='linear', method='2d')
 * of misfit/misfit.py:11-104, i.e. the back-and-forth optimal-transport solver misfit/QW2D/src/fot2d.c
 * (compute_l2_fot2d :514-606, fotGradient2d :608-656) that the reference runs per shot as a subprocess over files
 * (misfit/bfm.py:145-193), with the direct-wave subtraction of fwi.py:146-150 (dw nullable).
 * syn, obs, dw, adjsrc_out: [nshots][nt][nrec] fp32 (the solver's n1 = nrec, n2 = nt). Per record: positivity shift
 * c = gamma*max(0, -min), densities normalised to unit mean, `num_steps` back-and-forth iterations with initial step
 * step_scale / max density (marmousi2_fwi.py:131-132: gamma=1.01, num_steps=15, step_scale=4);
 * adjsrc_out = (dual - <mu, dual>) / mean(f) / mass;  fval_out[0] += sum over records of the W2 value;
 * loss_out (nullable): [nshots] per-record values. scratch: b2fwi_qw2d_scratch_bytes() bytes. Uses cuFFT (DCTs of
 * the Poisson solves); its plans are cached inside the library per record shape.
 */
int b2fwi_qw2d_misfit(const float *syn, const float *obs, const float *dw, int32_t nt, int32_t nrec, int32_t nshots,
                      double gamma, int32_t num_steps, float step_scale, float *adjsrc_out, double *fval_out,
                      float *loss_out, void *scratch, void *stream);
int64_t b2fwi_qw2d_scratch_bytes(int32_t nt, int32_t nrec, int32_t nshots);
/* One step of the solver on caller-provided single-record fields [nt][nrec] (diagnostics and the stage-by-stage
 * parity tests): op 0 out = c-transform(a) (fot2d.c:157-183); op 1 a += sigma * Poisson(b - c), scal[0] = H^-1
 * residual (fot2d.c:479-503); op 2 out = push-forward of density b by grad a (fot2d.c:290-478); op 3 scal[0] = W2
 * value of (phi a, dual b, mu c, nu d) (fot2d.c:519-531). scratch: b2fwi_qw2d_scratch_bytes(nt, nrec, 1). */
int b2fwi_qw2d_debug_step(int32_t op, int32_t nt, int32_t nrec, float *a, float *b, float *c, float *d, float sigma,
                          float *out, float *scal, void *scratch, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* B2FWI_H */

// stream_kernels.cuh -- launch wrappers of the streaming engine (see stream_kernels.cu)
#pragma once
#include "common.cuh"

namespace b2fwi {

struct StepArgs {
    int np, nr, nz, halo;
    int fs;                // free surface at z = 0: mirrored z stencil in the top rows (stream_point.cuh)
    int64_t sp, sr;
    float *out;            // slice written: u[t+1] (forward) or v[t-1] (backward)
    const float *cur;      // u[t] / v[t]
    const float *prev;     // u[t-1] / v[t+1]
    const float *c1, *c2;  // update coefficients (b2fwi_prepare_coeffs)
    const int *box;        // {lo_p, hi_p, lo_r, hi_r, lo_z, hi_z, valid}: inside, c1 == 1 exactly and is not read
    // imaging condition (IMG != 0): grad += -u.dt2 * cur
    float *grad;
    const float *h0, *h1, *h2;  // IMG==1: u[t-1], u[t], u[t+1];  IMG==2: h1 = u.dt2[t]
    int hist_uv;           // IMG==2 only: h1 = u[t] and the imaging factor is v.dt2[t] (B2FWI_HIST_UVDT2)
    float inv_dt2;
    // forward extras (nullable)
    float *illum;          // += cur^2
    float *d2u;            // = (prev - 2 cur + out) / dt^2
    // Laplacian weights pre-divided by h^2 for the plane / row / z directions
    float c0, c0_lo;       // centre weight as hi + lo: exactly -2*sum of the rounded side weights (see api.cu)
    float cp[B2FWI_MAX_R + 1], cr[B2FWI_MAX_R + 1], cz[B2FWI_MAX_R + 1];
    int chunk;             // planes per CTA along the streamed axis (3-D); <= 0: pick automatically
    // ---- sparse operators fused into the TMA sweeps (stream_tma.cu, service warps); all nullable
    b2fwi_sparse inj, itp;               // (by value) inj.row_tile != 0: inject `inj_vals` into `out`; itp.row_tile != 0: record `cur`
    const float *inj_vals;               // [inj.npoint] time sample of every injected point
    const float *vp;                     // injection scale dt^2 vp^2 (operators.py:134,221)
    float dt;
    float *itp_out;                      // [itp.npoint]
};

void fill_stencil_weights(const Layout &L, StepArgs *a);
int pick_chunk(const Layout &L);
int launch_step(const Layout &L, StepArgs a, int img, cudaStream_t st);
// TMA-staged 3-D sweeps: forward, adjoint + imaging from u.dt2 (stream_tma.cu)
bool tma_step_supported(const Layout &L, const StepArgs &a, int img);
// the sweep kernel can take this map's injection / interpolation on board (tables present, 3-D, TMA path)
bool tma_fusable(const Layout &L, const b2fwi_sparse *m, int what);   // what: 1 injection, 2 interpolation
void set_fuse(int mask);          // b2fwi_set_option("fuse", mask): 1 source injection, 2 interpolation, 4 any injection
int get_fuse();
int launch_step_tma(const Layout &L, const StepArgs &a, int img, cudaStream_t st);
void tma_tile_shape(int R, int *tz, int *tr);
bool tma_enabled(int img);
void set_tma_mask(int mask);      // b2fwi_set_option("tma", mask): bit 0 forward, bit 1 adjoint + imaging
int get_tma_mask();
// d2u/cur/prev: patch u.dt2 at the injected cells (forward); grad/hist: imaging by parts (B2FWI_HIST_UVDT2) - the
// sweep formed v.dt2 before the injection, the injected increment is added here: grad -= hist * dv / dt^2
int launch_inject(float *field, const float *vp, float dt, const float *vals, const b2fwi_sparse *m,
                  float *d2u, const float *cur, const float *prev, float inv_dt2, cudaStream_t st,
                  float *grad = nullptr, const float *hist = nullptr);
int launch_interp(const float *field, float *out, const b2fwi_sparse *m, cudaStream_t st);
// kernel='OT4': out += c2 * dt^2/12 * L(vp^2 L(cur)) through the scratch slice `tmp` (two launches)
int launch_ot4_correction(const Layout &L, const StepArgs &a, const float *vp, float dt, float *tmp, cudaStream_t st);
int launch_coeffs(const Layout &L, const float *vp, const float *damp, float dt, float *coef, cudaStream_t st);
int launch_accum_sq(const Layout &L, float *acc, const float *f, cudaStream_t st);
int launch_born_source(const Layout &L, float *field, const float *c2, const float *dm, const float *d2u,
                       cudaStream_t st);
int launch_geometry_mask(const b2fwi_grid *g, int nbl, const double *pts, int npts, double *mask, cudaStream_t st);
int launch_crop_mask_acc(const b2fwi_grid *g, const Layout &L, int nbl, const float *field, const double *mask,
                         double *out, cudaStream_t st);

}  // namespace b2fwi

// resident2d.cuh -- argument block of the SM-resident 2-D engine (see resident2d.cu)
#pragma once
#include "common.cuh"

namespace b2fwi {

#define RES2D_MAX_CELLS 1024   // injection cells per CTA (shared-memory staging, double buffered)
#define RES2D_MAX_CON 1024     // injection contributions (point, weight) per CTA, staged in shared memory
// 4-row-strip variant (resident2d_lat.cu): staged source / residual row (points used by one CTA) and receiver tables
#define RES2D_LAT_MAXV 1024
#define RES2D_LAT_MAXITP 1024

struct Res2dArgs {
    // ---- decomposition (b2fwi_res2d_plan)
    int nx, nz;            // padded grid (rows, contiguous z)
    int nzq;               // quads (float4) per row = ceil(nz / 4)
    int C;                 // CTAs per cluster == CTAs per shot; rows are split across them
    int rows_cta;          // rows owned by every CTA but the last = ceil(nx / C)
    int G;                 // row groups per CTA: thread (g, q) owns rows g*P .. g*P+P-1 of quad q
    int threads;
    int tile_rows;         // G*P + 2R
    // ---- window (imaging / history / illumination), padded coordinates; z in quads
    int wx0, wx1, wq0, wq1;
    // ---- time
    int nt, time_m, time_M;
    float inv_dt2;
    // ---- model (global memory)
    const float *B;        // dt^2 vp^2 as a pitched slice (row stride sr); zero in the pitch padding
    int64_t sr;
    const float *sx;       // [nx]     x part of damp/dt   (damp is separable: model.py:31-49)
    const float *sz;       // [nzq*4]  z part of damp/dt, zero padded
    float c0, c0_lo, cx[5], cz[5];
    // ---- shots
    int nshots;
    const float *vals;     // forward: src[shot][nt][nvals]; backward: residual[shot][nt][nvals]
    int nvals;
    int64_t vals_shot_stride;
    // injection maps, indexed by shot*C + cluster rank
    const int32_t *inj_desc;    // [.][2]: ncell, cell_base
    const int32_t *inj_cptr;    // CSR pointers of the cells into inj_pt / inj_w
    const int32_t *inj_pt;
    const float *inj_w;
    const unsigned long long *thr_mask;  // [.][threads] bit r*4+j: lane j of the thread's row r is an injection cell
    const int32_t *thr_base;             // [.][threads] first cell slot of the thread
    // receiver recording (forward)
    const int32_t *itp_desc;    // [.][2]: count, base
    const int32_t *itp_pt;      // receiver index
    const int32_t *itp_off;     // [.][4] offsets (floats) into the CTA tile, -1 = outside
    const float *itp_w;         // [.][4]
    float *rec;                 // [shot][nt][nrec]
    int nrec;
    // ---- wavefield history of the window: u.dt2, [shot][t - hist_t0][wx1-wx0][(wq1-wq0)*4]
    float *hist;
    int hist_t0;
    int64_t hist_shot_stride, hist_t_stride;
    // ---- window accumulator written at the end: illumination (forward) or gradient (backward)
    float *out;                 // [shot][wx1-wx0][(wq1-wq0)*4]
};

// short strips (3 or 4 rows per thread), few shots per GPU (resident2d_lat.cu)
#define RES2D_LAT_MAXP 4
int launch_res2d_lat(const Res2dArgs &a, int R, int P, int mode, cudaStream_t st);
int res2d_lat_pitch_quads(int nzq);      // tile row pitch of the short-strip kernels (0: rows too long for them)
size_t res2d_lat_smem_bytes(const Res2dArgs &a);
int res2d_lat_max_clusters(const Res2dArgs &a, int R, int P, int *out);

}  // namespace b2fwi

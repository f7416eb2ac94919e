// resident2d_lat.cu -- entry points of the short-strip (3 or 4 rows per thread) resident kernels; the kernels are
// instantiated per stencil radius in resident2d_lat_r{2,3,4}.cu (parallel compilation)
#include "resident2d_lat.cuh"

namespace b2fwi {

int lat_dispatch_r2(const Res2dArgs &a, int P, int mode, int op, cudaStream_t st, int *out);
int lat_dispatch_r3(const Res2dArgs &a, int P, int mode, int op, cudaStream_t st, int *out);
int lat_dispatch_r4(const Res2dArgs &a, int P, int mode, int op, cudaStream_t st, int *out);

static int dispatch(const Res2dArgs &a, int R, int P, int mode, int op, cudaStream_t st, int *out)
{
    switch (R) {
    case 2: return lat_dispatch_r2(a, P, mode, op, st, out);
    case 3: return lat_dispatch_r3(a, P, mode, op, st, out);
    case 4: return lat_dispatch_r4(a, P, mode, op, st, out);
    default: set_error("res2d: space order %d not supported by the resident engine", 2 * R); return B2FWI_EUNSUPPORTED;
    }
}

int res2d_lat_pitch_quads(int nzq) { return lat::pitch_quads(nzq); }

size_t res2d_lat_smem_bytes(const Res2dArgs &a)
{
    return lat::carve(a.tile_rows, a.nzq, a.rows_cta, (a.wq1 - a.wq0) * 4).total;
}

int launch_res2d_lat(const Res2dArgs &a, int R, int P, int mode, cudaStream_t st)
{
    return dispatch(a, R, P, mode, 0, st, nullptr);
}

int res2d_lat_max_clusters(const Res2dArgs &a, int R, int P, int *out) { return dispatch(a, R, P, 1, 1, nullptr, out); }

}  // namespace b2fwi

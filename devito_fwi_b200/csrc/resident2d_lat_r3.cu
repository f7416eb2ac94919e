// resident2d_lat_r3.cu -- instantiates the short-strip resident kernels (resident2d_lat.cuh) for space_order 6
#include "resident2d_lat.cuh"

namespace b2fwi {
int lat_dispatch_r3(const Res2dArgs &a, int P, int mode, int op, cudaStream_t st, int *out)
{
    return lat_dispatch<3>(a, P, mode, op, st, out);
}
}  // namespace b2fwi

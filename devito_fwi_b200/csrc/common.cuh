// common.cuh -- shared host/device helpers of libb2fwi (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "b2fwi.h"

#define B2FWI_MAX_R 8

namespace b2fwi {

void set_error(const char *fmt, ...);
void count_launch(int n = 1);   // bookkeeping behind b2fwi_launch_count()

#define B2_CHECK_ARG(cond, ...)              \
    do {                                     \
        if (!(cond)) {                       \
            b2fwi::set_error(__VA_ARGS__);   \
            return B2FWI_EINVAL;             \
        }                                    \
    } while (0)

#define B2_CUDA(call)                                                                        \
    do {                                                                                     \
        cudaError_t e_ = (call);                                                             \
        if (e_ != cudaSuccess) {                                                             \
            b2fwi::set_error("%s:%d: %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
            return B2FWI_ECUDA;                                                              \
        }                                                                                    \
    } while (0)

// Internal view of a grid: (plane, row, z). 3-D: (x, y, z). 2-D: one plane, rows = x.
struct Layout {
    int ndim;
    int np, nr, nz;      // domain extents; np == 1 in 2-D
    int halo;
    int64_t sp, sr;      // element strides of plane / row (sp unused in 2-D)
    int64_t base;        // offset of domain cell (0,0,0)
    int64_t elems;       // floats per haloed slice
    int R;               // stencil radius
    int fs;              // free surface at z index 0 (b2fwi_grid.fs)
    int ot4;             // fourth-order-in-time update (b2fwi_grid.kernel == 1)
    double inv_h2[3];    // 1/h^2 for plane, row, z directions
};

int make_layout(const b2fwi_grid *g, Layout *L);

// central second-derivative weights c[0..R] on a unit grid
void laplace_coeffs(int R, double *c);

}  // namespace b2fwi

// resident2d.cu -- SM-resident 2-D engine: one thread-block CLUSTER per shot, the whole time loop in
// one launch, wavefields never leave the chip.
//
// The named 2-D configurations (circle 281^2, Marmousi 380x186, Marmousi2 420x220) have 70-92 k grid
// points: one wavefield is ~300 KB, so a time step is bound by launch / synchronisation latency, not by
// HBM. This engine therefore keeps, per shot, on the C SMs of a cluster:
//   * u[t] in shared memory, double buffered, rows split across the cluster's CTAs; the R boundary rows
//     are pushed into the neighbour CTA's halo rows through distributed shared memory (st.shared::cluster)
//     and one barrier.cluster per time step orders everything;
//   * per grid point, in REGISTERS: delta = u[t] - u[t-1] (the update is carried in increment form) and
//     B = dt^2 vp^2; each thread owns a strip of P rows x 4 contiguous z and streams a 2R+1-row register
//     window down its strip (2.5-D register streaming along the slow axis, 128-bit shared loads along z);
//   * the sponge factor 1/(1 + dt*damp*vp^2) is evaluated on the fly from the separable damping profile;
//   * source / residual injection through a cell-centric gather staged in shared memory one step ahead,
//     receiver interpolation straight from the shared tile;
//   * the zero-lag imaging condition (backward) and the source illumination (forward) accumulate in a
//     shared-memory tile of the imaging window; only u.dt2 of that window streams to / from HBM.
// Update (same algebra as operators.py:87 / acoustic_time_update_nb.ipynb cell 3, written for delta):
//   delta+ = c1 * (delta + B * L(u)),  u+ = u + delta+,  c1 = 1/(1 + (damp/dt) * B),  c2 = B * c1.
// All shots of a rank run concurrently (grid = nshots clusters); arithmetic uses packed fp32x2 FMAs.
#include <cooperative_groups.h>

#include "common.cuh"
#include "resident2d.cuh"

namespace cg = cooperative_groups;

namespace b2fwi {

// ---- packed fp32x2 helpers (FFMA2 / FADD2 / FMUL2 on sm_100a)
static __device__ __forceinline__ float4 ld4s(const float *p) { return *reinterpret_cast<const float4 *>(p); }
static __device__ __forceinline__ void st4s(float *p, float4 v) { *reinterpret_cast<float4 *>(p) = v; }
static __device__ __forceinline__ float4 z4() { return make_float4(0.f, 0.f, 0.f, 0.f); }
static __device__ __forceinline__ float2 lo2(float4 a) { return make_float2(a.x, a.y); }
static __device__ __forceinline__ float2 hi2(float4 a) { return make_float2(a.z, a.w); }
static __device__ __forceinline__ float4 mk4(float2 l, float2 h) { return make_float4(l.x, l.y, h.x, h.y); }
static __device__ __forceinline__ float4 add4(float4 a, float4 b)
{
    return mk4(__fadd2_rn(lo2(a), lo2(b)), __fadd2_rn(hi2(a), hi2(b)));
}
static __device__ __forceinline__ float4 fma4s(float s, float4 a, float4 c)   // s*a + c
{
    const float2 ss = make_float2(s, s);
    return mk4(__ffma2_rn(ss, lo2(a), lo2(c)), __ffma2_rn(ss, hi2(a), hi2(c)));
}
static __device__ __forceinline__ float4 fma4(float4 a, float4 b, float4 c)   // a*b + c
{
    return mk4(__ffma2_rn(lo2(a), lo2(b), lo2(c)), __ffma2_rn(hi2(a), hi2(b), hi2(c)));
}
static __device__ __forceinline__ float4 mul4(float4 a, float4 b)
{
    return mk4(__fmul2_rn(lo2(a), lo2(b)), __fmul2_rn(hi2(a), hi2(b)));
}
static __device__ __forceinline__ float4 mul4s(float s, float4 a)
{
    const float2 ss = make_float2(s, s);
    return mk4(__fmul2_rn(ss, lo2(a)), __fmul2_rn(ss, hi2(a)));
}
static __device__ __forceinline__ float rcp_approx(float x)
{
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

template <int P>
struct MaxThreads { static constexpr int value = (P <= 8) ? 480 : (P <= 12) ? 384 : 288; };

// MODE 0: forward (record receivers, store u.dt2 history, accumulate illumination)
// MODE 1: backward with imaging condition (read u.dt2 history, accumulate gradient)
template <int R, int P, int MODE>
__global__ void __launch_bounds__(MaxThreads<P>::value, 1) res2d_kernel(const __grid_constant__ Res2dArgs a)
{
    extern __shared__ __align__(16) float smem[];
    cg::cluster_group cluster = cg::this_cluster();
    const int crank = (int)cluster.block_rank();
    const int shot = blockIdx.x / a.C;
    const int sc = shot * a.C + crank;
    const int tid = threadIdx.x;
    const int T = a.threads;
    const int pitch = a.nzq * 4;
    const int xg = tid / a.nzq, qi = tid - xg * a.nzq;
    const bool tactive = xg < a.G;
    const int row0 = crank * a.rows_cta;
    const int rows_valid = min(a.rows_cta, a.nx - row0);
    const int lr0 = xg * P;

    // ---- shared memory carve-up
    const int tile_elems = a.tile_rows * pitch;
    const int wcols = (a.wq1 - a.wq0) * 4;
    const int acc_lr0 = max(a.wx0 - row0, 0);                                   // first local row inside the window
    const int acc_rows = max(min(a.wx1 - row0, rows_valid) - acc_lr0, 0);
    float *tile0 = smem;
    float *tile1 = tile0 + tile_elems;
    float *acc = tile1 + tile_elems;                                            // [rows_cta][wcols]
    float *sxs = acc + a.rows_cta * wcols;                                      // [G*P]
    float *injb = sxs + a.G * P;                                                // [2][RES2D_MAX_CELLS]

    for (int i = tid; i < 2 * tile_elems + a.rows_cta * wcols; i += blockDim.x) smem[i] = 0.f;
    for (int i = tid; i < a.G * P; i += blockDim.x) sxs[i] = (row0 + i < a.nx && i < rows_valid) ? a.sx[row0 + i] : 0.f;

    // ---- per-thread persistent state
    float4 dl[P], Bq[P];
#pragma unroll
    for (int r = 0; r < P; r++) {
        dl[r] = z4();
        const bool ok = tactive && (lr0 + r < rows_valid);
        Bq[r] = ok ? __ldg(reinterpret_cast<const float4 *>(a.B + (int64_t)(row0 + lr0 + r) * a.sr + 4 * qi)) : z4();
    }
    const float4 szq = tactive ? __ldg(reinterpret_cast<const float4 *>(a.sz + 4 * qi)) : z4();
    const bool qin = tactive && qi >= a.wq0 && qi < a.wq1;
    const unsigned long long imask = tactive ? a.thr_mask[(int64_t)sc * T + tid] : 0ull;
    const int ibase = tactive ? a.thr_base[(int64_t)sc * T + tid] : 0;

    // injection descriptors of this CTA
    const int ncell = a.inj_desc[2 * sc], cell_base = a.inj_desc[2 * sc + 1];
    const float *vals = a.vals + (int64_t)shot * a.vals_shot_stride;
    auto gather = [&](int t, int s) -> float {
        float v = 0.f;
        const int j0 = a.inj_cptr[cell_base + s], j1 = a.inj_cptr[cell_base + s + 1];
        for (int j = j0; j < j1; j++) v = fmaf(a.inj_w[j], __ldg(vals + (int64_t)t * a.nvals + a.inj_pt[j]), v);
        return v;
    };

    // remote views of the neighbours' tiles (distributed shared memory)
    float *prev0 = (crank > 0) ? cluster.map_shared_rank(tile0, crank - 1) : nullptr;
    float *prev1 = (crank > 0) ? cluster.map_shared_rank(tile1, crank - 1) : nullptr;
    float *next0 = (crank < a.C - 1) ? cluster.map_shared_rank(tile0, crank + 1) : nullptr;
    float *next1 = (crank < a.C - 1) ? cluster.map_shared_rank(tile1, crank + 1) : nullptr;

    const int nsteps = a.time_M - a.time_m + 1;
    const int t_first = (MODE == 0) ? a.time_m : a.time_M;
    for (int s = tid; s < ncell; s += blockDim.x) injb[s] = gather(t_first, s);
    cluster.sync();

    float *hist = a.hist ? a.hist + (int64_t)shot * a.hist_shot_stride : nullptr;
    const int64_t hq = (int64_t)(a.wq1 - a.wq0) * 4;     // history / out row stride

    int cur = 0;
    for (int step = 0; step < nsteps; ++step) {
        const int t = (MODE == 0) ? a.time_m + step : a.time_M - step;
        const float *tc = cur ? tile1 : tile0;
        float *tn = cur ? tile0 : tile1;
        float *tprev = cur ? prev0 : prev1;
        float *tnext = cur ? next0 : next1;
        const float *injc = injb + (step & 1) * RES2D_MAX_CELLS;
        float *injn = injb + ((step + 1) & 1) * RES2D_MAX_CELLS;

        // injection values of the NEXT step: loads in flight while this step computes
        const bool more = step + 1 < nsteps;
        const int t_next = (MODE == 0) ? t + 1 : t - 1;
        float injv = 0.f;
        if (more && tid < ncell) injv = gather(t_next, tid);

        if (MODE == 0 && a.rec) {
            // rec[t][p] = sum_c w_c u[t][c]   (operators.py:137)
            const int cnt = a.itp_desc[2 * sc], base = a.itp_desc[2 * sc + 1];
            for (int i = tid; i < cnt; i += blockDim.x) {
                float sum = 0.f;
#pragma unroll
                for (int c = 0; c < 4; c++) {
                    const int off = a.itp_off[4 * (base + i) + c];
                    if (off >= 0) sum += a.itp_w[4 * (base + i) + c] * tc[off];
                }
                a.rec[((int64_t)shot * a.nt + t) * a.nrec + a.itp_pt[base + i]] = sum;
            }
        }

        if (tactive) {
            const float *hbase = hist ? hist + (int64_t)(t - a.hist_t0) * a.hist_t_stride : nullptr;
            float4 w[2 * R + 1];
#pragma unroll
            for (int i = 0; i < 2 * R; i++) w[i] = ld4s(tc + (lr0 + i) * pitch + 4 * qi);
            // history prefetch (backward): two rows ahead
            float4 hpre[2] = {z4(), z4()};
            if (MODE == 1) {
#pragma unroll
                for (int i = 0; i < 2; i++) {
                    const int lr = lr0 + i, row = row0 + lr;
                    if (i < P && qin && lr < rows_valid && row >= a.wx0 && row < a.wx1)
                        hpre[i] = __ldg(reinterpret_cast<const float4 *>(hbase + (int64_t)(row - a.wx0) * hq + 4 * (qi - a.wq0)));
                }
            }
#pragma unroll
            for (int r = 0; r < P; r++) {
                const int lr = lr0 + r;
                w[2 * R] = ld4s(tc + (lr + 2 * R) * pitch + 4 * qi);
                float4 hnow = hpre[0];
                if (MODE == 1) {
                    hpre[0] = hpre[1];
                    hpre[1] = z4();
                    const int lr2 = lr + 2, row2 = row0 + lr2;
                    if (r + 2 < P && qin && lr2 < rows_valid && row2 >= a.wx0 && row2 < a.wx1)
                        hpre[1] = __ldg(reinterpret_cast<const float4 *>(hbase + (int64_t)(row2 - a.wx0) * hq + 4 * (qi - a.wq0)));
                }
                if (lr < rows_valid) {
                    const float4 Cq = w[R];
                    const float *crow = tc + (lr + R) * pitch + 4 * qi;
                    const float4 Lq = (qi > 0) ? ld4s(crow - 4) : z4();
                    const float4 Rq = (qi + 1 < a.nzq) ? ld4s(crow + 4) : z4();
                    // Laplacian: centre (hi + lo weight), rows (register window), z (shared neighbours)
                    float4 lap = fma4s(a.c0, Cq, mul4s(a.c0_lo, Cq));
#pragma unroll
                    for (int k = 1; k <= R; k++) lap = fma4s(a.cx[k], add4(w[R + k], w[R - k]), lap);
                    const float zl[12] = {Lq.x, Lq.y, Lq.z, Lq.w, Cq.x, Cq.y, Cq.z, Cq.w, Rq.x, Rq.y, Rq.z, Rq.w};
                    float2 l01 = lo2(lap), l23 = hi2(lap);
#pragma unroll
                    for (int k = 1; k <= R; k++) {
                        const float2 ck = make_float2(a.cz[k], a.cz[k]);
                        l01 = __ffma2_rn(ck, __fadd2_rn(make_float2(zl[4 + k], zl[5 + k]), make_float2(zl[4 - k], zl[5 - k])), l01);
                        l23 = __ffma2_rn(ck, __fadd2_rn(make_float2(zl[6 + k], zl[7 + k]), make_float2(zl[6 - k], zl[7 - k])), l23);
                    }
                    lap = mk4(l01, l23);
                    // update in increment form
                    const float4 tmp = fma4(Bq[r], lap, dl[r]);
                    const float sxr = sxs[lr];
                    const float4 den = fma4(make_float4(sxr + szq.x, sxr + szq.y, sxr + szq.z, sxr + szq.w), Bq[r],
                                            make_float4(1.f, 1.f, 1.f, 1.f));
                    const float4 c1 = make_float4(rcp_approx(den.x), rcp_approx(den.y), rcp_approx(den.z), rcp_approx(den.w));
                    float4 dn = mul4(c1, tmp);
                    const unsigned rowbits = (unsigned)((imask >> (4 * r)) & 0xFull);
                    if (rowbits) {
                        const unsigned long long below = imask & ((1ull << (4 * r)) - 1ull);
                        int slot = ibase + __popcll(below);
                        if (rowbits & 1u) dn.x = fmaf(injc[slot++], Bq[r].x, dn.x);
                        if (rowbits & 2u) dn.y = fmaf(injc[slot++], Bq[r].y, dn.y);
                        if (rowbits & 4u) dn.z = fmaf(injc[slot++], Bq[r].z, dn.z);
                        if (rowbits & 8u) dn.w = fmaf(injc[slot++], Bq[r].w, dn.w);
                    }
                    const float4 un = add4(Cq, dn);
                    st4s(tn + (lr + R) * pitch + 4 * qi, un);
                    if (lr < R && tprev) st4s(tprev + (R + a.rows_cta + lr) * pitch + 4 * qi, un);
                    if (lr >= rows_valid - R && tnext) st4s(tnext + (lr - (rows_valid - R)) * pitch + 4 * qi, un);

                    const int row = row0 + lr;
                    if (qin && row >= a.wx0 && row < a.wx1) {
                        float *ap = acc + (lr - acc_lr0) * wcols + 4 * (qi - a.wq0);
                        if (MODE == 0) {
                            if (hbase) {
                                const float4 d2 = mul4s(a.inv_dt2, make_float4(dn.x - dl[r].x, dn.y - dl[r].y,
                                                                              dn.z - dl[r].z, dn.w - dl[r].w));
                                st4s(const_cast<float *>(hbase) + (int64_t)(row - a.wx0) * hq + 4 * (qi - a.wq0), d2);
                            }
                            if (a.out) st4s(ap, fma4(un, un, ld4s(ap)));            // illum += u[t+1]^2
                        } else {
                            // grad += -u.dt2[t] * v[t]   (operators.py:217)
                            st4s(ap, fma4(make_float4(-hnow.x, -hnow.y, -hnow.z, -hnow.w), Cq, ld4s(ap)));
                        }
                    }
                    dl[r] = dn;
                }
#pragma unroll
                for (int i = 0; i < 2 * R; i++) w[i] = w[i + 1];
            }
        }
        if (more) {
            if (tid < ncell) injn[tid] = injv;
            for (int s = tid + blockDim.x; s < ncell; s += blockDim.x) injn[s] = gather(t_next, s);
        }
        cluster.sync();
        cur ^= 1;
    }

    // ---- window accumulator -> global
    if (a.out) {
        float *out = a.out + (int64_t)shot * (int64_t)(a.wx1 - a.wx0) * hq;
        for (int i = tid; i < acc_rows * wcols; i += blockDim.x) {
            const int rr = i / wcols, cc = i - rr * wcols;
            out[(int64_t)(row0 + acc_lr0 + rr - a.wx0) * hq + cc] = acc[rr * wcols + cc];
        }
    }
}

// ------------------------------------------------------------------------------------------------
size_t res2d_smem_bytes(const Res2dArgs &a, int P)
{
    const size_t pitch = (size_t)a.nzq * 4;
    const size_t wcols = (size_t)(a.wq1 - a.wq0) * 4;
    return sizeof(float) * (2 * (size_t)a.tile_rows * pitch + (size_t)a.rows_cta * wcols + (size_t)a.G * P +
                            2 * RES2D_MAX_CELLS);
}

template <int R, int P, int MODE>
static int launch_one(const Res2dArgs &a, cudaStream_t st)
{
    const size_t smem = res2d_smem_bytes(a, P);
    auto kern = res2d_kernel<R, P, MODE>;
    B2_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3((unsigned)(a.nshots * a.C), 1, 1);
    cfg.blockDim = dim3((unsigned)a.threads, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)a.C;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    B2_CUDA(cudaLaunchKernelEx(&cfg, kern, a));
    count_launch();
    return 0;
}

template <int R, int MODE>
static int launch_P(const Res2dArgs &a, int P, cudaStream_t st)
{
    switch (P) {
    case 8: return launch_one<R, 8, MODE>(a, st);
    case 12: return launch_one<R, 12, MODE>(a, st);
    case 16: return launch_one<R, 16, MODE>(a, st);
    default: set_error("res2d: unsupported rows-per-thread %d", P); return B2FWI_EUNSUPPORTED;
    }
}

int launch_res2d(const Res2dArgs &a, int R, int P, int mode, cudaStream_t st)
{
    if (a.threads > (P <= 8 ? 480 : P <= 12 ? 384 : 288)) {
        set_error("res2d: %d threads exceed the limit for P=%d", a.threads, P);
        return B2FWI_EINVAL;
    }
#define B2_RCASE(r)                                                         \
    case r:                                                                 \
        return mode == 0 ? launch_P<r, 0>(a, P, st) : launch_P<r, 1>(a, P, st);
    switch (R) {
        B2_RCASE(2) B2_RCASE(3) B2_RCASE(4)
    default: set_error("res2d: space order %d not supported by the resident engine", 2 * R); return B2FWI_EUNSUPPORTED;
    }
#undef B2_RCASE
}

}  // namespace b2fwi

// resident2d_lat.cuh -- the SM-resident 2-D engine for FEW shots per GPU (strong scaling, line searches, single-shot
// calls): same decomposition, arithmetic and results as resident2d.cu (bit for bit), but built for the shortest time
// step instead of the fewest SMs per shot.
//
// With 4 shots on a 148-SM GPU a shot can have 16 SMs (a non-portable cluster size), i.e. ~24 grid rows per CTA.
// resident2d.cu at that size is bound by exposed latencies, not by issue slots (measured, Marmousi, 1 shot on 16 SMs:
// 2.4 us per step of which ~1.7 us do not depend on the strip length): B = dt^2 vp^2 and the u.dt2 history are re-read
// from L2 every row with a two-row prefetch distance that a 4-row strip cannot cover, the injection gather and the
// receiver maps chain through global memory in the few warps that own them, and everyone waits for those warps at the
// step barrier. Here a thread owns 4 rows x 4 z of one strip and
//   * keeps B and the sponge factor c1 = 1/(1 + (damp/dt) B) of its 16 points in REGISTERS (no per-step loads);
//   * reads the u.dt2 history (backward) from shared memory: one cp.async.bulk per step and CTA copies the CTA's slab
//     of the window, two time levels ahead, into a 3-deep ring and signals an mbarrier;
//   * reads injection values from a shared-memory copy of the source / residual row (4-byte cp.async two steps ahead),
//     so the cell gather is shared memory -> shared memory;
//   * keeps the receiver interpolation tables in shared memory and its own injection bitmap in a register.
// Halo rows travel exactly as in resident2d.cu (st.async + mbarrier complete_tx in the receiving CTA).
#pragma once
#include <cooperative_groups.h>
#include <string.h>

#include "common.cuh"
#include "packed.cuh"
#include "resident2d.cuh"

namespace cg = cooperative_groups;

namespace b2fwi {

namespace lat {

constexpr int NHB = 3;          // history ring depth (slabs in flight: 2)

static __device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
static __device__ __forceinline__ float4 lds4(uint32_t addr)
{
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}
static __device__ __forceinline__ int4 lds4i(uint32_t addr)
{
    int4 v;
    asm volatile("ld.shared.v4.s32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}
static __device__ __forceinline__ float lds1(uint32_t addr)
{
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
    return v;
}
static __device__ __forceinline__ void sts4(uint32_t addr, float4 v)
{
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
static __device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t rank)
{
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
static __device__ __forceinline__ void st_async4(uint32_t remote_addr, float4 v, uint32_t remote_bar)
{
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.f32 [%0], {%1, %2, %3, %4}, [%5];"
                 ::"r"(remote_addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "r"(remote_bar) : "memory");
}
static __device__ __forceinline__ void st_async4_if(unsigned pred, uint32_t remote_addr, float4 v, uint32_t remote_bar)
{
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %6, 0;\n\t"
                 "@q st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.f32 [%0], {%1, %2, %3, %4}, [%5];\n\t}"
                 ::"r"(remote_addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "r"(remote_bar), "r"(pred) : "memory");
}
static __device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
static __device__ __forceinline__ void mbar_arm(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
static __device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    uint32_t done, spins = 0;
    do {
        asm volatile("{\n\t.reg .pred p;\n\t"
                     "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                     "selp.b32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(bar), "r"(parity) : "memory");
        if (!done && ++spins > (1u << 26)) __trap();     // a lost halo / slab would otherwise hang the GPU
    } while (!done);
}
// global -> shared bulk copy (TMA engine, 1-D), completion counted in bytes on an mbarrier of this CTA
static __device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
static __device__ __forceinline__ void cp_async4(uint32_t dst, const void *src)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory");
}
static __device__ __forceinline__ void cp_async_wait_all()
{
    asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
}

struct Smem {                   // offsets in bytes
    uint32_t tile0, tile1, acc, injb, cw, cp, vrow, aux, bars, total;
};

// tile row pitch in quads: a compile-time constant of the kernel (row addresses become immediates)
__host__ __device__ inline int pitch_quads(int nzq) { return (nzq + 2 <= 64) ? 64 : (nzq + 2 <= 96) ? 96 : 0; }

__host__ __device__ inline Smem carve(int tile_rows, int nzq, int rows_cta, int wcols)
{
    Smem s;
    const uint32_t pitch = (uint32_t)pitch_quads(nzq) * 4u;
    const uint32_t tile = (uint32_t)tile_rows * pitch * 4u;
    uint32_t o = 0;
    s.tile0 = o; o += tile;
    s.tile1 = o; o += tile;
    s.acc = o; o += (uint32_t)rows_cta * (uint32_t)wcols * 4u;
    s.injb = o; o += 2u * RES2D_MAX_CELLS * 4u;
    s.cw = o; o += RES2D_MAX_CON * 4u;
    s.cp = o; o += RES2D_MAX_CON * 2u;
    s.vrow = o; o += 2u * RES2D_LAT_MAXV * 4u;
    // forward: receiver tables (off[4], w[4], pt) ; backward: history ring. One region, sized for the larger.
    const uint32_t itp = RES2D_LAT_MAXITP * 36u;
    const uint32_t hb = (uint32_t)NHB * (uint32_t)rows_cta * (uint32_t)wcols * 4u;
    s.aux = o; o += (itp > hb ? itp : hb);
    o = (o + 15u) & ~15u;
    s.bars = o; o += (2u + NHB) * 8u;
    s.total = o;
    return s;
}

}  // namespace lat

using namespace lat;

// MODE_ 0: forward (history / illumination optional), 1: backward + imaging, 2: forward of a gradient evaluation
// TMAX: launch bound (registers per thread: 168 with 3 warps per scheduler = 384 threads, 128 with 4 = 512)
template <int R, int P, int MODE_, int TMAX, int PQ>
__global__ void __launch_bounds__(TMAX, 1) res2d_lat_kernel(const __grid_constant__ Res2dArgs a)
{
    constexpr int MODE = (MODE_ == 2) ? 0 : MODE_;
    constexpr bool SAVE = (MODE_ == 2);
    constexpr bool C1REG = (TMAX <= 384);          // sponge factor in registers when the budget allows, else re-evaluated
    constexpr int NB = (TMAX <= 384) ? P : (P == 4 ? 2 : P);      // rows whose arithmetic is interleaved (register budget)
    static_assert(P == 3 || P == 4, "strips of 3 or 4 rows");
    extern __shared__ __align__(128) unsigned char smem_raw[];
    cg::cluster_group cluster = cg::this_cluster();
    const int crank = (int)cluster.block_rank();
    const int shot = blockIdx.x / a.C;
    const int sc = shot * a.C + crank;
    const int tid = threadIdx.x;
    const int T = a.threads;
    constexpr uint32_t pitchB = (uint32_t)PQ * 16u;      // tile row pitch in bytes
    const int xg = tid / a.nzq, qi = tid - xg * a.nzq;
    const bool tactive = xg < a.G;
    const int row0 = crank * a.rows_cta;
    const int rows_valid = min(a.rows_cta, a.nx - row0);
    const int lr0 = xg * P;
    const int wcols = (a.wq1 - a.wq0) * 4;
    const int acc_lr0 = max(a.wx0 - row0, 0);
    const int acc_rows = max(min(a.wx1 - row0, rows_valid) - acc_lr0, 0);

    const Smem so = carve(a.tile_rows, a.nzq, a.rows_cta, wcols);
    const uint32_t sbase = smem_u32(smem_raw);
    float *smf = reinterpret_cast<float *>(smem_raw);
    float *acc = reinterpret_cast<float *>(smem_raw + so.acc);
    float *injb = reinterpret_cast<float *>(smem_raw + so.injb);
    float *cw_s = reinterpret_cast<float *>(smem_raw + so.cw);
    unsigned short *cp_s = reinterpret_cast<unsigned short *>(smem_raw + so.cp);
    const uint32_t hbar = sbase + so.bars;           // [2] halo bytes of step parity
    const uint32_t hfull = hbar + 16u;               // [NHB] history slabs

    for (uint32_t i = tid; i < (so.injb) / 4u; i += blockDim.x) smf[i] = 0.f;      // tiles + accumulator

    // ---- per-thread persistent state: delta, B, c1 of the 4 x 4 points; per-row flags (4 bits per row:
    // 1 valid, 2 inside the imaging window, 4 push to previous CTA, 8 push to next CTA)
    float4 dl[P], Bq[P], c1q[P];
    float sxv[P];
    unsigned flags = 0u;
    const float4 szq = tactive ? __ldg(reinterpret_cast<const float4 *>(a.sz + 4 * qi)) : z4();
#pragma unroll
    for (int r = 0; r < P; r++) {
        dl[r] = z4();
        Bq[r] = z4();
        c1q[r] = z4();
        sxv[r] = 0.f;
        const int lr = lr0 + r, row = row0 + lr;
        if (tactive && lr < rows_valid) {
            unsigned f = 1u;
            if (qi >= a.wq0 && qi < a.wq1 && row >= a.wx0 && row < a.wx1) f |= 2u;
            if (lr < R && crank > 0) f |= 4u;
            if (lr >= rows_valid - R && crank < a.C - 1) f |= 8u;
            flags |= f << (4 * r);
            Bq[r] = __ldg(reinterpret_cast<const float4 *>(a.B + (int64_t)row * a.sr + 4 * qi));
            const float sxr = __ldg(a.sx + row);
            sxv[r] = sxr;
            if (C1REG) {
                const float4 den = fma4(add4(make_float4(sxr, sxr, sxr, sxr), szq), Bq[r], make_float4(1.f, 1.f, 1.f, 1.f));
                c1q[r] = make_float4(rcp_approx(den.x), rcp_approx(den.y), rcp_approx(den.z), rcp_approx(den.w));
            }
        }
    }
    // this thread's injection cells: bit r*4+j, and the first slot of the staged values
    unsigned imask = 0;
    int ibase = 0;
    if (tactive) {
        imask = (unsigned)(a.thr_mask[(int64_t)sc * T + tid] & ((1ull << (4 * P)) - 1ull));
        ibase = a.thr_base[(int64_t)sc * T + tid];
    }

    // ---- injection lists of this CTA -> shared memory; point indices relative to the first point used here
    const int ncell = a.inj_desc[2 * sc], cell_base = a.inj_desc[2 * sc + 1];
    const float *vals = a.vals + (int64_t)shot * a.vals_shot_stride;
    const int con0 = a.inj_cptr[cell_base];
    const int ncon = a.inj_cptr[cell_base + ncell] - con0;
    __shared__ int s_ptlo, s_pthi;
    if (tid == 0) { s_ptlo = 0x7fffffff; s_pthi = -1; }
    __syncthreads();
    {
        int lo = 0x7fffffff, hi = -1;
        for (int j = tid; j < ncon; j += blockDim.x) {
            const int p = a.inj_pt[con0 + j];
            lo = min(lo, p);
            hi = max(hi, p);
        }
        if (hi >= 0) { atomicMin(&s_ptlo, lo); atomicMax(&s_pthi, hi); }
    }
    __syncthreads();
    const int ptlo = (ncon > 0) ? s_ptlo : 0;
    const int npt = (ncon > 0) ? s_pthi - s_ptlo + 1 : 0;
    if (npt > RES2D_LAT_MAXV) __trap();              // the host plans around this (resident.build_maps)
    for (int j = tid; j < ncon; j += blockDim.x) {
        cw_s[j] = a.inj_w[con0 + j];
        cp_s[j] = (unsigned short)(a.inj_pt[con0 + j] - ptlo);
    }
    // contribution range of the cell slot this thread gathers every step (slot == stid)
    const int itp_cnt = (MODE == 0 && a.rec) ? a.itp_desc[2 * sc] : 0;
    const int itp_base = (MODE == 0 && a.rec) ? a.itp_desc[2 * sc + 1] : 0;
    if (itp_cnt > RES2D_LAT_MAXITP) __trap();
    const int nsvc = max(ncell, itp_cnt);
    const int svc0 = (max((int)blockDim.x - ((nsvc + 31) & ~31), 0) / 2) & ~31;        // service roles centred in the CTA
    const int stid = (tid >= svc0) ? tid - svc0 : tid + (int)blockDim.x - svc0;
    const bool needs_halo = (tactive && ((lr0 < R && crank > 0) || (lr0 + P - 1 + R >= rows_valid && crank < a.C - 1))) ||
                            (stid < itp_cnt);
    int gj0 = 0, gj1 = 0;
    if (stid < ncell) { gj0 = a.inj_cptr[cell_base + stid] - con0; gj1 = a.inj_cptr[cell_base + stid + 1] - con0; }
    // receiver tables (forward)
    const uint32_t ioff_s = sbase + so.aux, iw_s = ioff_s + RES2D_LAT_MAXITP * 16u, ipt_s = iw_s + RES2D_LAT_MAXITP * 16u;
    if (MODE == 0) {
        int *ioff = reinterpret_cast<int *>(smem_raw + so.aux);
        float *iw = reinterpret_cast<float *>(smem_raw + so.aux + RES2D_LAT_MAXITP * 16u);
        int *ipt = reinterpret_cast<int *>(smem_raw + so.aux + RES2D_LAT_MAXITP * 32u);
        for (int i = tid; i < itp_cnt; i += blockDim.x) {
#pragma unroll
            for (int c = 0; c < 4; c++) {
                ioff[4 * i + c] = a.itp_off[4 * (itp_base + i) + c];
                iw[4 * i + c] = a.itp_w[4 * (itp_base + i) + c];
            }
            ipt[i] = a.itp_pt[itp_base + i];
        }
    }

    // ---- 32-bit shared addresses
    const uint32_t own_off = (uint32_t)(lr0 + R) * pitchB + (uint32_t)(qi + 1) * 16u;
    uint32_t cur_s = sbase + so.tile0, nxt_s = sbase + so.tile1;
    uint32_t prv_n = 0, nex_n = 0, prv_c = 0, nex_c = 0;
    if (crank > 0) { prv_c = mapa_u32(cur_s, crank - 1); prv_n = mapa_u32(nxt_s, crank - 1); }
    if (crank < a.C - 1) { nex_c = mapa_u32(cur_s, crank + 1); nex_n = mapa_u32(nxt_s, crank + 1); }
    const uint32_t prev_delta = (uint32_t)a.rows_cta * pitchB;
    const uint32_t next_delta = (uint32_t)rows_valid * pitchB;
    const uint32_t accB = (uint32_t)wcols * 4u;
    const uint32_t win_off = (uint32_t)((lr0 - acc_lr0) * wcols + 4 * (qi - a.wq0)) * 4u;     // thread's row 0 in a window slab
    const uint32_t acc_s0 = sbase + so.acc + win_off;
    const uint32_t inj_s = sbase + so.injb;
    const uint32_t vrow_s = sbase + so.vrow;
    const uint32_t hb_s = sbase + so.aux;
    const uint32_t slabB = (uint32_t)(acc_rows * wcols) * 4u;

    const int nsteps = a.time_M - a.time_m + 1;
    const int t_first = (MODE == 0) ? a.time_m : a.time_M;
    const int tdir = (MODE == 0) ? 1 : -1;
    const int rows_next = min(a.rows_cta, a.nx - (crank + 1) * a.rows_cta);
    const uint32_t halo_bytes = (uint32_t)(((crank > 0 ? R : 0) + (crank < a.C - 1 ? min(R, max(rows_next, 0)) : 0)) *
                                           a.nzq * 16);
    uint32_t prv_bar = 0, nex_bar = 0;
    if (crank > 0) prv_bar = mapa_u32(hbar, crank - 1);
    if (crank < a.C - 1) nex_bar = mapa_u32(hbar, crank + 1);
    if (tid == 0) {
        mbar_init(hbar, 1);
        mbar_init(hbar + 8, 1);
#pragma unroll
        for (int b = 0; b < NHB; b++) mbar_init(hfull + 8u * b, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();      // lists staged, barriers initialised

    const int64_t hq = (int64_t)wcols;
    const float *hslab0 = nullptr;         // this CTA's slab of the window at the first time level (backward)
    const bool has_hist = SAVE || a.hist != nullptr;
    if (a.hist)
        hslab0 = a.hist + (int64_t)shot * a.hist_shot_stride + (int64_t)(t_first - a.hist_t0) * a.hist_t_stride +
                 (int64_t)(row0 + acc_lr0 - a.wx0) * hq;
    if (MODE == 1 && tid == 0 && slabB > 0) {
        // history slabs of the first two steps
        for (int k = 0; k < 2 && k < nsteps; k++) {
            mbar_arm(hfull + 8u * k, slabB);
            bulk_g2s(hb_s + (uint32_t)k * slabB, hslab0 - (int64_t)k * a.hist_t_stride, slabB, hfull + 8u * k);
        }
    }
    // injection values of the first step straight from global memory, the second step's row into vrow[1]
    for (int s = tid; s < ncell; s += blockDim.x) {
        float v = 0.f;
        const int j0 = a.inj_cptr[cell_base + s] - con0, j1 = a.inj_cptr[cell_base + s + 1] - con0;
        for (int j = j0; j < j1; j++) v = fmaf(cw_s[j], __ldg(vals + (int64_t)t_first * a.nvals + ptlo + cp_s[j]), v);
        injb[s] = v;
    }
    if (nsteps > 1)
        for (int i = tid; i < npt; i += blockDim.x)
            cp_async4(vrow_s + (uint32_t)(RES2D_LAT_MAXV + i) * 4u, vals + (int64_t)(t_first + tdir) * a.nvals + ptlo + i);
    cp_async_wait_all();
    cluster.sync();

    // forward history store: float4 index of this thread's row 0 at the current time level
    float4 *hbase = reinterpret_cast<float4 *>(a.hist);
    uint32_t hidx0 = 0;
    if (MODE == 0 && a.hist)
        hidx0 = (uint32_t)(((int64_t)shot * a.hist_shot_stride + (int64_t)(t_first - a.hist_t0) * a.hist_t_stride +
                            (int64_t)(row0 + lr0 - a.wx0) * hq + 4 * (qi - a.wq0)) / 4);
    const uint32_t hq4 = (uint32_t)(a.wq1 - a.wq0);
    const uint32_t hstep4 = (uint32_t)(a.hist_t_stride / 4);
    const float c0 = a.c0, c0_lo = a.c0_lo, inv_dt2 = a.inv_dt2;

    for (int step = 0; step < nsteps; ++step) {
        const int t = t_first + tdir * step;
        const uint32_t injc = inj_s + (uint32_t)(step & 1) * (RES2D_MAX_CELLS * 4u);
        float *injn = injb + ((step + 1) & 1) * RES2D_MAX_CELLS;
        const bool more = step + 1 < nsteps;

        if (tid == 0) {
            mbar_arm(hbar + 8u * (step & 1), halo_bytes);
            if (MODE == 1 && slabB > 0 && step + 2 < nsteps) {
                // u.dt2 slab two time levels ahead (its ring slot was last read in step - 1)
                const uint32_t b = (uint32_t)((step + 2) % NHB);
                mbar_arm(hfull + 8u * b, slabB);
                bulk_g2s(hb_s + b * slabB, hslab0 - (int64_t)(step + 2) * a.hist_t_stride, slabB, hfull + 8u * b);
            }
        }
        // source / residual row two steps ahead -> vrow[step & 1] (that buffer held the row of THIS step, consumed by
        // the gather of step - 1)
        if (step + 2 < nsteps) {
            const float *vsrc = vals + (int64_t)(t + 2 * tdir) * a.nvals + ptlo;
            for (int i = tid; i < npt; i += blockDim.x)
                cp_async4(vrow_s + (uint32_t)((step & 1) * RES2D_LAT_MAXV + i) * 4u, vsrc + i);
        }
        // injection values of the NEXT step: shared -> shared
        if (more && stid < ncell) {
            const uint32_t vr = vrow_s + (uint32_t)(((step + 1) & 1) * RES2D_LAT_MAXV) * 4u;
            float v = 0.f;
            for (int j = gj0; j < gj1; j++) v = fmaf(cw_s[j], lds1(vr + 4u * cp_s[j]), v);
            injn[stid] = v;
        }
        if (more)
            for (int s = stid + blockDim.x; s < ncell; s += blockDim.x) {
                const uint32_t vr = vrow_s + (uint32_t)(((step + 1) & 1) * RES2D_LAT_MAXV) * 4u;
                const int j0 = a.inj_cptr[cell_base + s] - con0, j1 = a.inj_cptr[cell_base + s + 1] - con0;
                float v = 0.f;
                for (int j = j0; j < j1; j++) v = fmaf(cw_s[j], lds1(vr + 4u * cp_s[j]), v);
                injn[s] = v;
            }

        // the neighbours' boundary rows of the previous step: only the threads that read halo rows wait for them (the
        // strips next to the CTA's edges and the receiver-recording threads); every other warp starts its rows right
        // away, so the flight time of the pushes and the skew between neighbouring CTAs hide behind interior work
        if (needs_halo && step > 0) mbar_wait(hbar + 8u * ((step - 1) & 1), (uint32_t)((step - 1) >> 1) & 1u);
        if (MODE == 0 && a.rec) {
            // rec[t][p] = sum_c w_c u[t][c]   (operators.py:137)
            for (int i = stid; i < itp_cnt; i += blockDim.x) {
                const int4 o = lds4i(ioff_s + 16u * i);
                const float4 w = lds4(iw_s + 16u * i);
                float sum = 0.f;
                if (o.x >= 0) sum += w.x * lds1(cur_s + (uint32_t)o.x * 4u);
                if (o.y >= 0) sum += w.y * lds1(cur_s + (uint32_t)o.y * 4u);
                if (o.z >= 0) sum += w.z * lds1(cur_s + (uint32_t)o.z * 4u);
                if (o.w >= 0) sum += w.w * lds1(cur_s + (uint32_t)o.w * 4u);
                int pt;
                asm volatile("ld.shared.s32 %0, [%1];" : "=r"(pt) : "r"(ipt_s + 4u * i));
                a.rec[((int64_t)shot * a.nt + t) * a.nrec + pt] = sum;
            }
        }
        uint32_t hs = 0;
        if (MODE == 1 && has_hist) {
            const uint32_t b = (uint32_t)(step % NHB);
            if (slabB > 0) mbar_wait(hfull + 8u * b, (uint32_t)(step / NHB) & 1u);
            hs = hb_s + b * slabB + win_off;
        }

        if (tactive) {
            // Rows are processed in batches of NB: (A) every shared-memory load of the batch is issued first, (B) the
            // arithmetic of the NB rows is independent straight-line code the compiler interleaves (the row update is
            // a ~150-cycle dependent chain, and a CTA of this kernel has only 2-4 warps per scheduler to hide it),
            // (C) stores and halo pushes, (D) window accumulators. Rows past the end of the grid are computed like
            // any other (their B and c1 are zero, the tile rows they read stay zero) and only their stores are
            // predicated off: no branches around the arithmetic.
            constexpr int NWR = P + 2 * R;
            float4 w[NWR];                                             // rows lr0 - R .. lr0 + P - 1 + R of u[t]
            {
                const uint32_t rw = cur_s + own_off - (uint32_t)R * pitchB;
#pragma unroll
                for (int i = 0; i < NWR; i++) w[i] = lds4(rw + (uint32_t)i * pitchB);
            }
            const uint32_t ro0 = cur_s + own_off, rn0 = nxt_s + own_off;
#pragma unroll
            for (int rb = 0; rb < P; rb += NB) {
                float4 Lq[NB], Rq[NB], un[NB], dn[NB];
#pragma unroll
                for (int j = 0; j < NB; j++) {
                    Lq[j] = lds4(ro0 + (uint32_t)(rb + j) * pitchB - 16u);
                    Rq[j] = lds4(ro0 + (uint32_t)(rb + j) * pitchB + 16u);
                }
#pragma unroll
                for (int j = 0; j < NB; j++) {
                    const int r = rb + j;
                    const float4 Cq = w[r + R];
                    float4 lx = fma4s(c0, Cq, mul4s(c0_lo, Cq));
#pragma unroll
                    for (int k = 1; k <= R; k++) lx = fma4s(a.cx[k], add4(w[r + R + k], w[r + R - k]), lx);
                    const float zl[12] = {Lq[j].x, Lq[j].y, Lq[j].z, Lq[j].w, Cq.x, Cq.y, Cq.z, Cq.w,
                                          Rq[j].x, Rq[j].y, Rq[j].z, Rq[j].w};
                    // z neighbours z[i+k] + z[i-k], i = 4..7. Even k: both operands are aligned register pairs (FADD2);
                    // odd k: four scalar adds written straight into aligned pairs (a packed add would first need
                    // two register moves per operand). Same IEEE operations either way.
                    float2 l01, l23;
                    {
                        const float2 s01 = make_float2(__fadd_rn(zl[5], zl[3]), __fadd_rn(zl[6], zl[4]));
                        const float2 s23 = make_float2(__fadd_rn(zl[7], zl[5]), __fadd_rn(zl[8], zl[6]));
                        const float2 c1k = make_float2(a.cz[1], a.cz[1]);
                        l01 = __fmul2_rn(c1k, s01);
                        l23 = __fmul2_rn(c1k, s23);
                    }
#pragma unroll
                    for (int k = 2; k <= R; k++) {
                        const float2 ck = make_float2(a.cz[k], a.cz[k]);
                        float2 s01, s23;
                        if (k & 1) {
                            s01 = make_float2(__fadd_rn(zl[4 + k], zl[4 - k]), __fadd_rn(zl[5 + k], zl[5 - k]));
                            s23 = make_float2(__fadd_rn(zl[6 + k], zl[6 - k]), __fadd_rn(zl[7 + k], zl[7 - k]));
                        } else {
                            s01 = __fadd2_rn(make_float2(zl[4 + k], zl[5 + k]), make_float2(zl[4 - k], zl[5 - k]));
                            s23 = __fadd2_rn(make_float2(zl[6 + k], zl[7 + k]), make_float2(zl[6 - k], zl[7 - k]));
                        }
                        l01 = __ffma2_rn(ck, s01, l01);
                        l23 = __ffma2_rn(ck, s23, l23);
                    }
                    const float4 lap = add4(lx, mk4(l01, l23));
                    const float4 tmp = fma4(Bq[r], lap, dl[r]);
                    float4 c1;
                    if (C1REG) {
                        c1 = c1q[r];
                    } else {
                        const float4 den = fma4(add4(make_float4(sxv[r], sxv[r], sxv[r], sxv[r]), szq), Bq[r],
                                                make_float4(1.f, 1.f, 1.f, 1.f));
                        c1 = make_float4(rcp_approx(den.x), rcp_approx(den.y), rcp_approx(den.z), rcp_approx(den.w));
                    }
                    dn[j] = mul4(c1, tmp);
                    un[j] = add4(Cq, dn[j]);
                }
                // injection cells of this thread (rare): patch the increment, redo the sum
                if ((imask >> (4 * rb)) & ((1u << (4 * NB)) - 1u)) {
#pragma unroll
                    for (int j = 0; j < NB; j++) {
                        const int r = rb + j;
                        const unsigned rowbits = (imask >> (4 * r)) & 0xFu;
                        if (rowbits) {
                            uint32_t sl = injc + 4u * (uint32_t)(ibase + __popc(imask & ((1u << (4 * r)) - 1u)));
                            if (rowbits & 1u) { dn[j].x = fmaf(lds1(sl), Bq[r].x, dn[j].x); sl += 4u; }
                            if (rowbits & 2u) { dn[j].y = fmaf(lds1(sl), Bq[r].y, dn[j].y); sl += 4u; }
                            if (rowbits & 4u) { dn[j].z = fmaf(lds1(sl), Bq[r].z, dn[j].z); sl += 4u; }
                            if (rowbits & 8u) { dn[j].w = fmaf(lds1(sl), Bq[r].w, dn[j].w); sl += 4u; }
                            un[j] = add4(w[r + R], dn[j]);
                        }
                    }
                }
#pragma unroll
                for (int j = 0; j < NB; j++) {
                    const int r = rb + j;
                    const uint32_t roff = (uint32_t)r * pitchB;
                    if (flags & (1u << (4 * r))) sts4(rn0 + roff, un[j]);
                    // boundary rows go to the neighbour's halo right away: their flight overlaps the rest of the step
                    st_async4_if(flags & (4u << (4 * r)), prv_n + own_off + roff + prev_delta, un[j], prv_bar + 8u * (step & 1));
                    st_async4_if(flags & (8u << (4 * r)), nex_n + own_off + roff - next_delta, un[j], nex_bar + 8u * (step & 1));
                    if (MODE == 0 && has_hist) {
                        // u.dt2[t] = (delta+ - delta) / dt^2; streaming store: written once, read much later
                        const float4 d2 = mul4s(inv_dt2, add4(dn[j], make_float4(-dl[r].x, -dl[r].y, -dl[r].z, -dl[r].w)));
                        if (flags & (2u << (4 * r))) __stcs(hbase + (hidx0 + (uint32_t)r * hq4), d2);
                    }
                    dl[r] = dn[j];
                }
                // window accumulators: illum += u[t+1]^2 (forward), grad += -u.dt2[t] * v[t] (backward, operators.py:217)
                if (MODE == 1 || SAVE || a.out) {
                    float4 av[NB], hv[NB];
#pragma unroll
                    for (int j = 0; j < NB; j++) {
                        const int r = rb + j;
                        av[j] = z4();
                        hv[j] = z4();
                        if (flags & (2u << (4 * r))) av[j] = lds4(acc_s0 + (uint32_t)r * accB);
                        if (MODE == 1 && (flags & (2u << (4 * r)))) hv[j] = lds4(hs + (uint32_t)r * accB);
                    }
#pragma unroll
                    for (int j = 0; j < NB; j++) {
                        const int r = rb + j;
                        const float4 v = (MODE == 1) ? fma4(make_float4(-hv[j].x, -hv[j].y, -hv[j].z, -hv[j].w), w[r + R], av[j])
                                                     : fma4(un[j], un[j], av[j]);
                        if (flags & (2u << (4 * r))) sts4(acc_s0 + (uint32_t)r * accB, v);
                    }
                }
            }
        }
        cp_async_wait_all();                                          // this thread's part of the staged value row
        __syncthreads();                                              // u[t+1] rows, staging buffers written
        { uint32_t x = cur_s; cur_s = nxt_s; nxt_s = x; }
        { uint32_t x = prv_c; prv_c = prv_n; prv_n = x; }
        { uint32_t x = nex_c; nex_c = nex_n; nex_n = x; }
        hidx0 += hstep4;
    }

    if (nsteps > 0) mbar_wait(hbar + 8u * ((nsteps - 1) & 1), (uint32_t)((nsteps - 1) >> 1) & 1u);   // last pushes have landed
    cluster.sync();       // no CTA leaves while a neighbour could still address its shared memory
    if (a.out) {
        float *out = a.out + (int64_t)shot * (int64_t)(a.wx1 - a.wx0) * hq;
        for (int i = tid; i < acc_rows * wcols; i += blockDim.x) {
            const int rr = i / wcols, cc = i - rr * wcols;
            out[(int64_t)(row0 + acc_lr0 + rr - a.wx0) * hq + cc] = acc[rr * wcols + cc];
        }
    }
}

// ------------------------------------------------------------------------------------------------
// launch plumbing; one translation unit per stencil radius instantiates LatLaunch<R> (resident2d_lat_r*.cu)
template <int R, int P, int MODE, int TMAX, int PQ>
static int lat_config(const Res2dArgs &a, cudaLaunchConfig_t *cfg, cudaLaunchAttribute *attr, unsigned nclusters,
                      void (**kern_out)(const Res2dArgs))
{
    const size_t smem = carve(a.tile_rows, a.nzq, a.rows_cta, (a.wq1 - a.wq0) * 4).total;
    auto kern = res2d_lat_kernel<R, P, MODE, TMAX, PQ>;
    B2_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (a.C > 8) B2_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    memset(cfg, 0, sizeof(*cfg));
    cfg->gridDim = dim3(nclusters * (unsigned)a.C, 1, 1);
    cfg->blockDim = dim3((unsigned)a.threads, 1, 1);
    cfg->dynamicSmemBytes = smem;
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)a.C;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg->attrs = attr;
    cfg->numAttrs = 1;
    *kern_out = kern;
    return 0;
}

// op 0: launch mode `mode` on `st`; op 1: cudaOccupancyMaxActiveClusters into *out
template <int R, int P, int MODE, int TMAX, int PQ>
static int lat_do(const Res2dArgs &a, int op, cudaStream_t st, int *out)
{
    cudaLaunchConfig_t cfg;
    cudaLaunchAttribute attr[1];
    void (*kern)(const Res2dArgs);
    int rc = lat_config<R, P, MODE, TMAX, PQ>(a, &cfg, attr, op == 0 ? (unsigned)a.nshots : 64u, &kern);
    if (rc) return rc;
    if (op == 1) {
        B2_CUDA(cudaOccupancyMaxActiveClusters(out, kern, &cfg));
        return 0;
    }
    cfg.stream = st;
    B2_CUDA(cudaLaunchKernelEx(&cfg, kern, a));
    count_launch();
    return 0;
}

template <int R, int P, int TMAX, int PQ>
static int lat_mode(const Res2dArgs &a, int P_unused, int mode, int op, cudaStream_t st, int *out)
{
    (void)P_unused;
    if (op == 1 || mode != 0) return lat_do<R, P, 1, TMAX, PQ>(a, op, st, out);
    return (a.hist && a.out) ? lat_do<R, P, 2, TMAX, PQ>(a, op, st, out) : lat_do<R, P, 0, TMAX, PQ>(a, op, st, out);
}

template <int R>
int lat_dispatch(const Res2dArgs &a, int P, int mode, int op, cudaStream_t st, int *out)
{
    const int pq = pitch_quads(a.nzq);
    if (pq == 0 || a.threads > 512 || (P != 3 && P != 4)) {
        set_error("res2d (short strips): unsupported plan (P=%d, %d threads, %d quads per row)", P, a.threads, a.nzq);
        return B2FWI_EUNSUPPORTED;
    }
    const bool small = a.threads <= 384;
#define B2_LAT(p, q)                                                                          \
    if (P == p && pq == q)                                                                    \
        return small ? lat_mode<R, p, 384, q>(a, P, mode, op, st, out) : lat_mode<R, p, 512, q>(a, P, mode, op, st, out);
    B2_LAT(3, 64) B2_LAT(3, 96) B2_LAT(4, 64) B2_LAT(4, 96)
#undef B2_LAT
    return B2FWI_EUNSUPPORTED;
}

}  // namespace b2fwi

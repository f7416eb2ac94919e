// resident2d.cu -- SM-resident 2-D engine: one thread-block CLUSTER per shot, the whole time loop in
// one launch, wavefields never leave the chip.
//
// The named 2-D configurations (circle 281^2, Marmousi 380x186, Marmousi2 420x220) have 70-92 k grid
// points: one wavefield is ~300 KB, so a time step is bound by launch / synchronisation latency, not by
// HBM. This engine therefore keeps, per shot, on the C SMs of a cluster:
//   * u[t] in shared memory, double buffered, rows split across the cluster's CTAs; the R boundary rows
//     are pushed into the neighbour CTA's halo rows through distributed shared memory as st.async stores that
//     complete a byte count on an mbarrier of the receiving CTA: per step a CTA waits for its own threads
//     (bar.sync) and for its neighbours' bytes, there is no cluster-wide barrier;
//   * per grid point, in REGISTERS: delta = u[t] - u[t-1] (the update is carried in increment form) and
//     B = dt^2 vp^2; each thread owns a strip of P rows x 4 contiguous z and streams a 2R+1-row register
//     window down its strip (2.5-D register streaming along the slow axis, 128-bit shared loads along z);
//   * the sponge factor 1/(1 + dt*damp*vp^2) is evaluated on the fly from the separable damping profile;
//   * source / residual injection through a cell-centric gather staged in shared memory one step ahead,
//     receiver interpolation straight from the shared tile;
//   * the zero-lag imaging condition (backward) and the source illumination (forward) accumulate in a
//     shared-memory tile of the imaging window; only u.dt2 of that window streams to / from HBM (backward: one
//     bulk L2 prefetch per step pulls the next time level's slab in ahead of the per-row loads).
// Update (same algebra as operators.py:87 / acoustic_time_update_nb.ipynb cell 3, written for delta):
//   delta+ = c1 * (delta + B * L(u)),  u+ = u + delta+,  c1 = 1/(1 + (damp/dt) * B),  c2 = B * c1.
// All shots of a rank run concurrently (grid = nshots clusters); arithmetic uses packed fp32x2 FMAs.
#include <cooperative_groups.h>

#include "common.cuh"
#include "packed.cuh"
#include "resident2d.cuh"

namespace cg = cooperative_groups;

namespace b2fwi {

// ---- raw shared / distributed-shared accessors on 32-bit addresses (no generic-address arithmetic in the hot loop)
static __device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
static __device__ __forceinline__ float4 lds4(uint32_t addr)
{
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
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
static __device__ __forceinline__ void sts4_cluster(uint32_t addr, float4 v)
{
    asm volatile("st.shared::cluster.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
                 : "memory");
}
static __device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t rank)
{
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
// streaming 128-bit global load that does not allocate an L1 line: with ~219 KB of the 228 KB L1/shared array
// used as shared memory only ~70 L1 lines are left, and allocating loads (B and history prefetches of 12 warps)
// would serialise on them
static __device__ __forceinline__ float4 ldg_stream4(const float4 *p)
{
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}
static __device__ __forceinline__ void cluster_barrier()
{
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// ---- halo exchange without a cluster-wide barrier (B2FWI_RES2D_ASYNC_HALO, default on)
// The boundary rows travel as st.async stores that complete a transaction count on an mbarrier in the RECEIVING
// CTA; a CTA then only waits for (a) its own threads (bar.sync) and (b) the bytes of its one or two neighbours.
// barrier.cluster.arrive.release made every warp drain ALL its outstanding memory operations first - including the
// streaming u.dt2 history stores to HBM (ncu: 8 % of the forward kernel in ERRBAR/membar, 10 % in the barrier wait).
#ifndef B2FWI_RES2D_ASYNC_HALO
#define B2FWI_RES2D_ASYNC_HALO 1
#endif
static __device__ __forceinline__ void st_async4(uint32_t remote_addr, float4 v, uint32_t remote_bar)
{
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.f32 [%0], {%1, %2, %3, %4}, [%5];"
                 ::"r"(remote_addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "r"(remote_bar) : "memory");
}
static __device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
static __device__ __forceinline__ void mbar_arm(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// (default .acquire.cta scope: the halo rows are written straight into this SM's shared memory before the byte count
// completes, and shared memory is not cached - a cluster-scope acquire would only add an L1 invalidation
// (CCTL.IVALL) per step, which throws away the L1-resident receiver maps)
static __device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity)
{
    uint32_t done, spins = 0;
    do {
        asm volatile("{\n\t.reg .pred p;\n\t"
                     "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                     "selp.b32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(bar), "r"(parity) : "memory");
        if (!done && ++spins > (1u << 26)) __trap();     // a lost halo would otherwise hang the GPU: fail loudly instead
    } while (!done);
}

// ---- split step barrier (B2FWI_RES2D_SPLIT_BARRIER, default on)
// bar.sync at the end of a step made every warp wait for the slowest one with nothing to do. The CTA-wide step barrier is
// an mbarrier instead: a warp ARRIVES when its rows of u[t+1] are written, then issues what the next step needs from
// global memory (B / history of its first rows, the gathered injection values, the L2 prefetches), and only WAITS
// right before it touches the shared tiles again - the load latencies and the warps' skew overlap.
#ifndef B2FWI_RES2D_SPLIT_BARRIER
#define B2FWI_RES2D_SPLIT_BARRIER 1
#endif
static __device__ __forceinline__ void mbar_arrive_cta(uint32_t bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

// Hide a value from the optimiser: stops ptxas/nvvm from strength-reducing the per-row addresses of the
// unrolled row loop into P separately carried registers (which spilled the register-resident state).
static __device__ __forceinline__ void opaque(uint32_t &x) { asm volatile("" : "+r"(x)); }
template <typename T>
static __device__ __forceinline__ void opaque_ptr(T *&p) { asm volatile("" : "+l"(p)); }

template <int P>
struct MaxThreads { static constexpr int value = (P <= 8) ? 480 : (P <= 12) ? 384 : 288; };

// MODE 0: forward (record receivers, store u.dt2 history, accumulate illumination)
// MODE 1: backward with imaging condition (read u.dt2 history, accumulate gradient)
// MODE 2: forward of a gradient evaluation - history AND illumination are known to be requested, which removes the two
//         loop-invariant pointer tests from every row (the compiler re-tests them per row to save predicate registers)
template <int R, int P, int MODE_>
__global__ void __launch_bounds__(MaxThreads<P>::value, 1) res2d_kernel(const __grid_constant__ Res2dArgs a)
{
    constexpr int MODE = (MODE_ == 2) ? 0 : MODE_;
    constexpr bool SAVE = (MODE_ == 2);
    // split step barrier: the backward sweep only (its warps wait 15 % of their time at bar.sync and have the history
    // loads to overlap; the forward sweeps lose 1 % to the extra instructions)
    constexpr bool SPLIT = B2FWI_RES2D_ASYNC_HALO && B2FWI_RES2D_SPLIT_BARRIER && (MODE_ == 1);
    extern __shared__ __align__(16) float smem[];
    cg::cluster_group cluster = cg::this_cluster();
    const int crank = (int)cluster.block_rank();
    const int shot = blockIdx.x / a.C;
    const int sc = shot * a.C + crank;
    const int tid = threadIdx.x;
    const int T = a.threads;
    const int pitch = (a.nzq + 2) * 4;      // one always-zero quad on each side: no edge branches in the z stencil
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
    float *cw_s = injb + 2 * RES2D_MAX_CELLS;                                   // [RES2D_MAX_CON] contribution weights
    unsigned short *cp_s = reinterpret_cast<unsigned short *>(cw_s + RES2D_MAX_CON);   // [RES2D_MAX_CON] point indices
    const uint32_t hbar = smem_u32(cp_s + RES2D_MAX_CON);                       // [2] mbarriers: halo bytes of step parity

    for (int i = tid; i < 2 * tile_elems + a.rows_cta * wcols; i += blockDim.x) smem[i] = 0.f;
    for (int i = tid; i < a.G * P; i += blockDim.x) sxs[i] = (row0 + i < a.nx && i < rows_valid) ? a.sx[row0 + i] : 0.f;

    // ---- per-thread persistent state
    // delta lives in registers; B = dt^2 vp^2 is re-read from L2 every step with a two-row prefetch
    // (holding it in registers as well spills: 2 x P float4 + the 2R+1-row window exceed the budget)
    float4 dl[P];
    // per-row flags, 4 bits per row: 1 valid row, 2 inside the imaging window, 4 push to previous CTA, 8 push to next CTA
    unsigned long long flags = 0ull;
    // global operands are addressed as (uniform base pointer) + 32-bit float4 index: one register per running
    // index instead of 64-bit pointers (the kernel is register-bound)
    const float4 *Bbase = reinterpret_cast<const float4 *>(a.B);
    const uint32_t Bidx0 = (uint32_t)(((int64_t)(row0 + lr0) * a.sr + 4 * qi) / 4);
    const uint32_t Bstride = (uint32_t)(a.sr / 4);           // row stride in float4
#pragma unroll
    for (int r = 0; r < P; r++) {
        dl[r] = z4();
        const int lr = lr0 + r, row = row0 + lr;
        const bool ok = tactive && (lr < rows_valid);
        if (ok) {
            unsigned f = 1u;
            if (qi >= a.wq0 && qi < a.wq1 && row >= a.wx0 && row < a.wx1) f |= 2u;
            if (lr < R && crank > 0) f |= 4u;
            if (lr >= rows_valid - R && crank < a.C - 1) f |= 8u;
            flags |= (unsigned long long)f << (4 * r);
        }
    }
    const float4 szq = tactive ? __ldg(reinterpret_cast<const float4 *>(a.sz + 4 * qi)) : z4();
    // rows pushed to the neighbours. Forward kernel: a compact mask (bit r: to the previous CTA, bit 16+r: to the
    // next), walked bit by bit; the backward kernel has no register to spare for it and tests the flags instead.
    unsigned pmask = 0;
    bool push_any = false;
#pragma unroll
    for (int r = 0; r < P; r++) {
        if ((flags >> (4 * r)) & 4ull) pmask |= 1u << r;
        if ((flags >> (4 * r)) & 8ull) pmask |= 0x10000u << r;
        push_any = push_any || (((flags >> (4 * r)) & 12ull) != 0ull);
    }
    // injection cells owned by this thread are rare: only the row bitmap lives in a register, the lane mask and the
    // first slot index are re-read from global memory (L1/L2) inside the rarely taken branch
    const unsigned long long *imask_p = a.thr_mask + (int64_t)sc * T + tid;
    const int *ibase_p = a.thr_base + (int64_t)sc * T + tid;
    unsigned injrows = 0;                     // bit r: row r of the strip holds injection cells
    if (tactive) {
        const unsigned long long im = *imask_p;
#pragma unroll
        for (int r = 0; r < P; r++)
            if ((im >> (4 * r)) & 0xFull) injrows |= 1u << r;
    }
    // injection descriptors of this CTA. The (point, weight) lists never change during the sweep: they are staged
    // in shared memory once, so that the per-step gather is ONE level of independent global loads (the time
    // sample of each contributing point) instead of a three-deep dependent chain through global index arrays -
    // which used to delay the gathering warps by ~1000 cycles per step and, through the barrier, everyone.
    const int ncell = a.inj_desc[2 * sc], cell_base = a.inj_desc[2 * sc + 1];
    const float *vals = a.vals + (int64_t)shot * a.vals_shot_stride;
    const int con0 = a.inj_cptr[cell_base];
    const int ncon = a.inj_cptr[cell_base + ncell] - con0;
    for (int j = tid; j < ncon; j += blockDim.x) {
        cw_s[j] = a.inj_w[con0 + j];
        cp_s[j] = (unsigned short)a.inj_pt[con0 + j];
    }
    auto gather = [&](int t, int s) -> float {
        float v = 0.f;
        const int j0 = a.inj_cptr[cell_base + s] - con0, j1 = a.inj_cptr[cell_base + s + 1] - con0;
        for (int j = j0; j < j1; j++) v = fmaf(cw_s[j], __ldg(vals + (int64_t)t * a.nvals + cp_s[j]), v);
        return v;
    };
    // contribution range of the slot this thread gathers every step (slot == tid)
    int gj0 = 0, gj1 = 0;
    // "service" roles (injection gather, receiver recording) are dealt out from the middle of the CTA: the first and
    // last warps own the boundary strips and already carry the halo pushes
    const int nsvc = (MODE == 0 && a.rec) ? max(ncell, a.itp_desc[2 * sc]) : ncell;      // busiest service role
    const int svc0 = (max((int)blockDim.x - ((nsvc + 31) & ~31), 0) / 2) & ~31;          // ... centred in the CTA
    const int stid = (tid >= svc0) ? tid - svc0 : tid + (int)blockDim.x - svc0;
    if (stid < ncell) { gj0 = a.inj_cptr[cell_base + stid] - con0; gj1 = a.inj_cptr[cell_base + stid + 1] - con0; }

    // ---- 32-bit shared addresses (bytes); everything below is an offset from these
    const uint32_t pitchB = (uint32_t)pitch * 4u;
    const uint32_t own_off = (uint32_t)(lr0 + R) * pitchB + (uint32_t)(qi + 1) * 16u;   // own row 0 inside a tile buffer
    uint32_t cur_s = smem_u32(tile0), nxt_s = smem_u32(tile1);
    // neighbours' tiles through distributed shared memory; constant row displacement (see header comment)
    uint32_t prv_c = 0, prv_n = 0, nex_c = 0, nex_n = 0;
    if (crank > 0) { prv_c = mapa_u32(cur_s, crank - 1); prv_n = mapa_u32(nxt_s, crank - 1); }
    if (crank < a.C - 1) { nex_c = mapa_u32(cur_s, crank + 1); nex_n = mapa_u32(nxt_s, crank + 1); }
    const uint32_t prev_delta = (uint32_t)a.rows_cta * pitchB;      // own row lr -> row R + rows_cta + lr of the previous CTA
    const uint32_t next_delta = (uint32_t)rows_valid * pitchB;      // own row lr -> row lr - (rows_valid - R) of the next CTA
    const uint32_t acc_s0 = smem_u32(acc) + (uint32_t)((lr0 - acc_lr0) * wcols + 4 * (qi - a.wq0)) * 4u;
    const uint32_t accB = (uint32_t)wcols * 4u;
    const uint32_t sxs_s = smem_u32(sxs) + (uint32_t)lr0 * 4u;
    const uint32_t inj_s = smem_u32(injb);

    const int nsteps = a.time_M - a.time_m + 1;
    const int t_first = (MODE == 0) ? a.time_m : a.time_M;
    // halo bytes this CTA receives per step: R rows from the previous CTA, min(R, its valid rows) from the next
    const int rows_next = min(a.rows_cta, a.nx - (crank + 1) * a.rows_cta);
    const uint32_t halo_bytes = (uint32_t)(((crank > 0 ? R : 0) + (crank < a.C - 1 ? min(R, max(rows_next, 0)) : 0)) *
                                           a.nzq * 16);
    uint32_t prv_bar = 0, nex_bar = 0;      // the neighbours' mbarrier pair, through distributed shared memory
    if (crank > 0) prv_bar = mapa_u32(hbar, crank - 1);
    if (crank < a.C - 1) nex_bar = mapa_u32(hbar, crank + 1);
    const uint32_t sbar = hbar + 16u;         // step barrier (one arrival per warp)
    if (B2FWI_RES2D_ASYNC_HALO && tid == 0) {
        mbar_init(hbar, 1);
        mbar_init(hbar + 8, 1);
        mbar_init(sbar, (blockDim.x + 31) / 32);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();      // cw_s / cp_s staged
    for (int s = tid; s < ncell; s += blockDim.x) injb[s] = gather(t_first, s);
    cluster.sync();
    if (SPLIT) {
        __syncwarp();
        if ((tid & 31) == 0) mbar_arrive_cta(sbar);   // "step -1 is written": the first wait below passes
    }

    // history pointer of this thread's first row at the first time level; advanced by +-one slice per step
    const int64_t hq = (int64_t)(a.wq1 - a.wq0) * 4;     // history / out row stride (floats)
    const uint32_t hq4 = (uint32_t)(a.wq1 - a.wq0);      // ... in float4
    float4 *hbase = reinterpret_cast<float4 *>(a.hist);   // whole history < 2^32 float4 (64 GB): 32-bit indices
    uint32_t hidx0 = 0;                                   // this thread's first row at the current time level
    if (a.hist)
        hidx0 = (uint32_t)(((int64_t)shot * a.hist_shot_stride + (int64_t)(t_first - a.hist_t0) * a.hist_t_stride +
                            (int64_t)(row0 + lr0 - a.wx0) * hq + 4 * (qi - a.wq0)) / 4);
    const uint32_t hstep4 = (uint32_t)(((MODE == 0) ? a.hist_t_stride : -a.hist_t_stride) / 4);   // wraps mod 2^32
    const float c0 = a.c0, c0_lo = a.c0_lo, inv_dt2 = a.inv_dt2;
    const bool has_hist = SAVE || a.hist != nullptr;

    for (int step = 0; step < nsteps; ++step) {
        const int t = (MODE == 0) ? a.time_m + step : a.time_M - step;
        const uint32_t injc = inj_s + (uint32_t)(step & 1) * (RES2D_MAX_CELLS * 4u);
        float *injn = injb + ((step + 1) & 1) * RES2D_MAX_CELLS;

        if (B2FWI_RES2D_ASYNC_HALO && tid == 0) mbar_arm(hbar + 8u * (step & 1), halo_bytes);
        if (MODE == 1 && has_hist && tid == 32 && step + 1 < nsteps && acc_rows > 0) {
            // the history is streamed from HBM (5 GB per sweep, each value read once): one bulk L2 prefetch per step
            // pulls this CTA's slab of the NEXT time level in, so the per-row loads below find it in L2
            const float *nxt = a.hist + (int64_t)shot * a.hist_shot_stride +
                               (int64_t)(t - 1 - a.hist_t0) * a.hist_t_stride + (int64_t)(row0 + acc_lr0 - a.wx0) * hq;
            asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(nxt), "r"((uint32_t)(acc_rows * hq * 4)) : "memory");
        }
        if (MODE == 1 && tid == 64 && crank == 0 && step + 2 < nsteps) {
            // same for the residual row gathered at the top of the NEXT step (47 MB of residuals do not stay in L2
            // next to the history stream): row t-2, widened to 16-byte bounds
            const uintptr_t p0 = (uintptr_t)(vals + (int64_t)(t - 2) * a.nvals);
            const uintptr_t lo = p0 & ~(uintptr_t)15, hi = (p0 + (uintptr_t)a.nvals * 4 + 15) & ~(uintptr_t)15;
            asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(lo), "r"((uint32_t)(hi - lo)) : "memory");
        }
        // injection values of the NEXT step: loads in flight while this step computes
        const bool more = step + 1 < nsteps;
        const int t_next = (MODE == 0) ? t + 1 : t - 1;
        float injv = 0.f;
        if (more) {
            const float *vrow = vals + (int64_t)t_next * a.nvals;
            for (int j = gj0; j < gj1; j++) injv = fmaf(cw_s[j], __ldg(vrow + cp_s[j]), injv);
        }

        // B / history of the first two rows: issued first, so that their L2 latency runs under the recording below
        // and the window fill instead of stalling row 0
        uint32_t hp = hidx0;
        uint32_t bp = Bidx0;
        opaque(bp);
        float4 bpre[2], hpre[2];
        if (tactive) {
            if (flags & 0x01ull) bpre[0] = ldg_stream4(Bbase + bp);
            if (P > 1 && (flags & 0x10ull)) bpre[1] = ldg_stream4(Bbase + (bp + Bstride));
            if (MODE == 1) {
                if (flags & 0x02ull) hpre[0] = ldg_stream4(hbase + hp);
                if (P > 1 && (flags & 0x20ull)) hpre[1] = ldg_stream4(hbase + (hp + hq4));
            }
        }

        if (SPLIT) {
            // every warp of this CTA has written its rows of u[t] (and the injection staging), and the neighbours'
            // boundary rows have landed
            mbar_wait_cluster(sbar, (uint32_t)step & 1u);
            if (step > 0) mbar_wait_cluster(hbar + 8u * ((step - 1) & 1), (uint32_t)((step - 1) >> 1) & 1u);
        }
        if (MODE == 0 && a.rec) {
            // rec[t][p] = sum_c w_c u[t][c]   (operators.py:137)
            const int cnt = a.itp_desc[2 * sc], base = a.itp_desc[2 * sc + 1];
            for (int i = stid; i < cnt; i += blockDim.x) {
                float sum = 0.f;
#pragma unroll
                for (int c = 0; c < 4; c++) {
                    const int off = a.itp_off[4 * (base + i) + c];
                    if (off >= 0) sum += a.itp_w[4 * (base + i) + c] * lds1(cur_s + (uint32_t)off * 4u);
                }
                a.rec[((int64_t)shot * a.nt + t) * a.nrec + a.itp_pt[base + i]] = sum;
            }
        }

        if (tactive) {
            // register window over rows lr-R .. lr+R, addressed circularly with compile-time slots (no moves)
            constexpr int NW = 2 * R + 1;
            float4 w[NW];
            uint32_t rw = cur_s + own_off - (uint32_t)R * pitchB;      // row lr0 - R
#pragma unroll
            for (int i = 0; i < 2 * R; i++) { w[i] = lds4(rw); rw += pitchB; }
            uint32_t ro = cur_s + own_off;                             // own row, current buffer
            uint32_t rn = nxt_s + own_off;                             // own row, next buffer
            uint32_t ra = acc_s0;
#pragma unroll
            for (int r = 0; r < P; r++) {
                w[(r + 2 * R) % NW] = lds4(rw);
                rw += pitchB;
                const unsigned f = (unsigned)(flags >> (4 * r)) & 0xFu;
                const unsigned f2 = (r + 2 < P) ? (unsigned)(flags >> (4 * ((r + 2) & 15))) & 0xFu : 0u;
                const float4 Bq = bpre[r & 1];
                if (f2 & 1u) bpre[r & 1] = ldg_stream4(Bbase + (bp + 2 * Bstride));
                float4 hnow;
                if (MODE == 1) {
                    hnow = hpre[r & 1];
                    if (f2 & 2u) hpre[r & 1] = ldg_stream4(hbase + (hp + 2 * hq4));
                }
                if (f & 1u) {
                    const float4 Cq = w[(r + R) % NW];
                    if (MODE == 1 && (f & 2u)) {
                        // grad += -u.dt2[t] * v[t]   (operators.py:217); done first so the history registers die early
                        sts4(ra, fma4(make_float4(-hnow.x, -hnow.y, -hnow.z, -hnow.w), Cq, lds4(ra)));
                    }
                    const float4 Lq = lds4(ro - 16u);
                    const float4 Rq = lds4(ro + 16u);
                    // Laplacian in two independent chains: centre (hi + lo weight) + rows from the register
                    // window, and the z neighbours from shared memory
                    float4 lx = fma4s(c0, Cq, mul4s(c0_lo, Cq));
#pragma unroll
                    for (int k = 1; k <= R; k++) lx = fma4s(a.cx[k], add4(w[(r + R + k) % NW], w[(r + R - k + NW) % NW]), lx);
                    const float zl[12] = {Lq.x, Lq.y, Lq.z, Lq.w, Cq.x, Cq.y, Cq.z, Cq.w, Rq.x, Rq.y, Rq.z, Rq.w};
                    // odd offsets pair up registers that are not aligned pairs (two MOVs per operand of a packed add):
                    // their sums are formed by scalar adds straight into an aligned pair instead (same values)
                    const float2 c1k = make_float2(a.cz[1], a.cz[1]);
                    float2 l01 = __fmul2_rn(c1k, make_float2(__fadd_rn(zl[5], zl[3]), __fadd_rn(zl[6], zl[4])));
                    float2 l23 = __fmul2_rn(c1k, make_float2(__fadd_rn(zl[7], zl[5]), __fadd_rn(zl[8], zl[6])));
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
                    // update in increment form
                    const float4 tmp = fma4(Bq, lap, dl[r]);
                    const float sxr = lds1(sxs_s + 4u * r);
                    const float4 den = fma4(add4(make_float4(sxr, sxr, sxr, sxr), szq), Bq, make_float4(1.f, 1.f, 1.f, 1.f));
                    const float4 c1 = make_float4(rcp_approx(den.x), rcp_approx(den.y), rcp_approx(den.z), rcp_approx(den.w));
                    float4 dn = mul4(c1, tmp);
                    if (injrows & (1u << r)) {
                        const unsigned long long imask = __ldg(imask_p);
                        const unsigned rowbits = (unsigned)((imask >> (4 * r)) & 0xFull);
                        const unsigned long long below = imask & ((1ull << (4 * r)) - 1ull);
                        uint32_t sl = injc + 4u * (uint32_t)(__ldg(ibase_p) + __popcll(below));
                        if (rowbits & 1u) { dn.x = fmaf(lds1(sl), Bq.x, dn.x); sl += 4u; }
                        if (rowbits & 2u) { dn.y = fmaf(lds1(sl), Bq.y, dn.y); sl += 4u; }
                        if (rowbits & 4u) { dn.z = fmaf(lds1(sl), Bq.z, dn.z); sl += 4u; }
                        if (rowbits & 8u) { dn.w = fmaf(lds1(sl), Bq.w, dn.w); sl += 4u; }
                    }
                    const float4 un = add4(Cq, dn);
                    sts4(rn, un);
                    if (MODE == 0 && (f & 2u)) {
                        {
                            if (has_hist) {
                                // u.dt2[t] = (delta+ - delta) / dt^2; streaming store: written once, read much later
                                const float4 d2 = mul4s(inv_dt2, add4(dn, make_float4(-dl[r].x, -dl[r].y, -dl[r].z, -dl[r].w)));
                                __stcs(hbase + hp, d2);
                            }
                            if (SAVE || a.out) sts4(ra, fma4(un, un, lds4(ra)));    // illum += u[t+1]^2
                        }
                    }
                    dl[r] = dn;
                }
                ro += pitchB;
                rn += pitchB;
                ra += accB;
                hp += hq4;
                bp += Bstride;
                opaque(rw); opaque(ro); opaque(rn); opaque(ra);
                opaque(hp); opaque(bp);
            }
        }
        // push this CTA's R boundary rows into the neighbours' halo rows (distributed shared memory). Done after the
        // row loop, by the few threads that own boundary rows, re-reading what they just wrote: keeps the ~10
        // address-conversion instructions of a remote store out of every row of every thread.
        if (MODE == 0) {
            for (unsigned m = pmask & 0xffffu; m; m &= m - 1) {
                const uint32_t off = own_off + (uint32_t)(__ffs(m) - 1) * pitchB;
                const float4 un = lds4(nxt_s + off);
#if B2FWI_RES2D_ASYNC_HALO
                st_async4(prv_n + off + prev_delta, un, prv_bar + 8u * (step & 1));
#else
                sts4_cluster(prv_n + off + prev_delta, un);
#endif
            }
            for (unsigned m = pmask >> 16; m; m &= m - 1) {
                const uint32_t off = own_off + (uint32_t)(__ffs(m) - 1) * pitchB;
                const float4 un = lds4(nxt_s + off);
#if B2FWI_RES2D_ASYNC_HALO
                st_async4(nex_n + off - next_delta, un, nex_bar + 8u * (step & 1));
#else
                sts4_cluster(nex_n + off - next_delta, un);
#endif
            }
        } else if (push_any) {
#pragma unroll
            for (int r = 0; r < P; r++) {
                const unsigned f = (unsigned)(flags >> (4 * r)) & 0xFu;
                if (f & 12u) {
                    const uint32_t off = own_off + (uint32_t)r * pitchB;
                    const float4 un = lds4(nxt_s + off);
#if B2FWI_RES2D_ASYNC_HALO
                    if (f & 4u) st_async4(prv_n + off + prev_delta, un, prv_bar + 8u * (step & 1));
                    if (f & 8u) st_async4(nex_n + off - next_delta, un, nex_bar + 8u * (step & 1));
#else
                    if (f & 4u) sts4_cluster(prv_n + off + prev_delta, un);
                    if (f & 8u) sts4_cluster(nex_n + off - next_delta, un);
#endif
                }
            }
        }
        if (more) {
            if (stid < ncell) injn[stid] = injv;
            for (int s = stid + blockDim.x; s < ncell; s += blockDim.x) injn[s] = gather(t_next, s);
        }
#if B2FWI_RES2D_ASYNC_HALO
        if (SPLIT) {
            __syncwarp();                                             // this warp's rows of u[t+1] (and its staging) are written
            if ((tid & 31) == 0) mbar_arrive_cta(sbar);
        } else {
            __syncthreads();                                          // this CTA's rows of u[t+1] (and the staging) are written
            mbar_wait_cluster(hbar + 8u * (step & 1), (uint32_t)(step >> 1) & 1u);   // ... and the neighbours' boundary rows have landed
        }
#else
        cluster_barrier();
#endif
        // swap buffers
        { uint32_t x = cur_s; cur_s = nxt_s; nxt_s = x; }
        { uint32_t x = prv_c; prv_c = prv_n; prv_n = x; }
        { uint32_t x = nex_c; nex_c = nex_n; nex_n = x; }
        hidx0 += hstep4;
    }

#if B2FWI_RES2D_ASYNC_HALO
    if (SPLIT && nsteps > 0)
        mbar_wait_cluster(hbar + 8u * ((nsteps - 1) & 1), (uint32_t)((nsteps - 1) >> 1) & 1u);   // last halo bytes
    cluster.sync();       // no CTA leaves while a neighbour could still address its shared memory
#endif
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
    if (P <= RES2D_LAT_MAXP) return res2d_lat_smem_bytes(a);
    const size_t pitch = ((size_t)a.nzq + 2) * 4;
    const size_t wcols = (size_t)(a.wq1 - a.wq0) * 4;
    return sizeof(float) * (2 * (size_t)a.tile_rows * pitch + (size_t)a.rows_cta * wcols + (size_t)a.G * P +
                            2 * RES2D_MAX_CELLS + RES2D_MAX_CON) + sizeof(unsigned short) * RES2D_MAX_CON +
           3 * sizeof(unsigned long long);
}

template <int R, int P, int MODE>
static int launch_one(const Res2dArgs &a, cudaStream_t st)
{
    const size_t smem = res2d_smem_bytes(a, P);
    auto kern = res2d_kernel<R, P, MODE>;
    B2_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (a.C > 8) B2_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));   // up to 16 SMs per shot
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

template <int R, int P>
static int max_clusters_one(const Res2dArgs &a, int *out)
{
    const size_t smem = res2d_smem_bytes(a, P);
    auto kern = res2d_kernel<R, P, 0>;
    B2_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (a.C > 8) B2_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3((unsigned)(a.C * 64), 1, 1);
    cfg.blockDim = dim3((unsigned)a.threads, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)a.C;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    B2_CUDA(cudaOccupancyMaxActiveClusters(out, kern, &cfg));
    return 0;
}

// how many clusters (shots) of this plan the device runs concurrently
int res2d_max_clusters(const Res2dArgs &a, int R, int P, int *out)
{
#define B2_OCASE(r, p) if (R == r && P == p) return max_clusters_one<r, p>(a, out);
    if (P <= RES2D_LAT_MAXP) return res2d_lat_max_clusters(a, R, P, out);
    B2_OCASE(2, 8) B2_OCASE(2, 12) B2_OCASE(2, 16) B2_OCASE(3, 8) B2_OCASE(3, 12) B2_OCASE(3, 16)
    B2_OCASE(4, 8) B2_OCASE(4, 12) B2_OCASE(4, 16)
#undef B2_OCASE
    set_error("res2d: unsupported (R, P) = (%d, %d)", R, P);
    return B2FWI_EUNSUPPORTED;
}

int launch_res2d(const Res2dArgs &a, int R, int P, int mode, cudaStream_t st)
{
    if (P <= RES2D_LAT_MAXP) return launch_res2d_lat(a, R, P, mode, st);
    if (a.threads > (P <= 4 ? 512 : P <= 8 ? 480 : P <= 12 ? 384 : 288)) {
        set_error("res2d: %d threads exceed the limit for P=%d", a.threads, P);
        return B2FWI_EINVAL;
    }
#define B2_RCASE(r)                                                         \
    case r:                                                                 \
        return mode != 0 ? launch_P<r, 1>(a, P, st)                         \
                         : (a.hist && a.out) ? launch_P<r, 2>(a, P, st) : launch_P<r, 0>(a, P, st);
    switch (R) {
        B2_RCASE(2) B2_RCASE(3) B2_RCASE(4)
    default: set_error("res2d: space order %d not supported by the resident engine", 2 * R); return B2FWI_EUNSUPPORTED;
    }
#undef B2_RCASE
}

}  // namespace b2fwi

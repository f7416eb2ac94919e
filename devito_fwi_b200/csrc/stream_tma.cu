// stream_tma.cu -- 3-D forward sweep of the streaming engine with TMA-staged tiles (sm_100a).
//
// Same sweep as step_kernel<R, 3, 0> (stream_kernels.cu) -- the OT2 update of
// seismic/acoustic/operators.py:59-95 -- with the data movement handed to the Tensor Memory Accelerator:
//
//   * a CTA owns a 128 z x 16 row tile and streams along the plane (x) axis; warps 0..15 compute (one float4 per
//     thread and plane, the plane-axis neighbours in a rotating register pipeline of 2R+1 float4), warp 16 is
//     the producer: one lane issues cp.async.bulk.tensor loads and never touches the data;
//   * u[t]: one box of (16 + 2R) rows x (128 + 8) z per plane (row / z halos included, zero-filled outside the
//     grid by the TMA unit - no predicates, no halo threads) into a ring of 2R+1 stages. A plane enters the SM
//     once: it feeds the register pipeline when it is R planes ahead and serves as the row/z neighbourhood tile
//     when its turn comes;
//   * u[t-1] and the two update coefficients: 128 x 16 boxes into a second ring (NPCC stages); inside the
//     undamped interior box (coef[0] == 1 exactly) the c1 load is skipped as in step_kernel;
//   * full/empty mbarriers per stage instead of a CTA-wide barrier per plane; every smem stage offset and
//     register slot is a compile-time constant (the plane loop is unrolled 2R+1 times).
//
// The arithmetic is point_update() of stream_point.cuh, so the sweep is bitwise identical to step_kernel's.
#include <cuda.h>
#include <stdlib.h>

#include <type_traits>

#include "common.cuh"
#include "packed.cuh"
#include "stream_kernels.cuh"
#include "stream_point.cuh"

namespace b2fwi {

// ---- mbarrier / TMA primitives (PTX ISA 8.x: mbarrier, cp.async.bulk.tensor)
static __device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
static __device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
static __device__ __forceinline__ void mbar_arrive(uint32_t bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
static __device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity)
{
    uint32_t done;
    asm volatile("{\n\t.reg .pred p;\n\t"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                 "selp.b32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    return done != 0;
}
static __device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    while (!mbar_try_wait(bar, parity)) { }
}
static __device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap *map, uint32_t bar,
                                                   int cz, int cr, int cp)
{
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes"
                 " [%0], [%1, {%3, %4, %5}], [%2];"
                 ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(cz), "r"(cr), "r"(cp)
                 : "memory");
}
static __device__ __forceinline__ float4 lds4(uint32_t addr)
{
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}

template <int R, int NPCC, int IMG>
struct TmaCfg {
    // so <= 8: 128 z x 16 rows, 16 compute warps. so > 8: the 2R+1 float4 register pipeline needs > 96 registers per
    // thread, so the tile is 64 z x 16 rows with 8 compute warps (register cap 224) and a 17-plane ring of
    // correspondingly smaller boxes still fits 227 KB.
    static constexpr int TZQ = (R <= 4) ? 32 : 16, TR = 16, TZ = 4 * TZQ, ZH = 4 * ((R + 3) / 4);
    static constexpr int SW = TZ + 2 * ZH, SROWS = TR + 2 * R;
    static constexpr int NQ = 2 * R + 1, NCUR = NQ;
    static constexpr int CUR_BYTES = SROWS * SW * 4;                       // one u[t] box
    static constexpr int CUR_STRIDE = (CUR_BYTES + 127) / 128 * 128;       // ring stage stride (TMA destinations: 128 B)
    static constexpr int PL_BYTES = TR * TZ * 4;                           // one pointwise-operand box
    static constexpr int NARR = (IMG >= 2) ? 4 : 3;                        // prev, c2, c1 [, u.dt2]
    static constexpr int PCC_BYTES = NARR * PL_BYTES;
    // operand-ring stage of unroll slot j is the compile-time j % NPCC when the ring turns a whole, odd number of
    // times per group of NQ planes; otherwise stage and parity are carried in two registers
    static constexpr bool PCT = (NQ % NPCC == 0) && (((NQ / NPCC) & 1) == 1);
    // warps are allocated in groups of four: 16 (8) compute warps + the producer leave three warps that cost no
    // registers. They are the SERVICE warps: receiver interpolation of the tile (from global memory, free running)
    // and the staging of the injection terms of plane i (warp i % NSVC, stage i % NI) that the compute threads add
    // to their result before they store it.
    static constexpr int NSVC = 3, NI = 6, NE = (R <= 4) ? 8 : 4;          // NE = B2FWI max_row_con the kernel takes
    static constexpr int NCONS = TZQ * TR, NTHREADS = NCONS + 32 + 32 * NSVC;
    static constexpr int BAR_OFF = NCUR * CUR_STRIDE + NPCC * PCC_BYTES;    // mbarriers (64 slots reserved)
    static constexpr int W_OFF = BAR_OFF + 64 * 8;                          // Laplacian weights, 32 floats
    static constexpr int TH_OFF = W_OFF + 32 * 4;                           // per-thread constants, uint4 each
    static constexpr int NMW = 32;                                          // plane-mask words: chunks of <= 1024 planes
    static constexpr int IMASK_OFF = TH_OFF + NCONS * 16;                   // [NMW] bit i: plane i has injection cells in this tile's rows
    static constexpr int ICNT_OFF = IMASK_OFF + NMW * 4;                    // [NI][TR] staged terms per row
    static constexpr int IENT_OFF = ICNT_OFF + NI * TR * 4;                 // [NI][TR][NE] {term, z in tile}
    static constexpr int SMEM = IENT_OFF + NI * TR * NE * 8;
    static_assert(R >= 1 && R <= 8, "stencil radius");
    static_assert(PL_BYTES % 128 == 0, "TMA destinations are 128-byte aligned");
    static_assert(2 * NCUR + NPCC + 2 * NI + 1 <= 64 && 2 + 3 * R <= 32, "barrier / weight slots");
};

// IMG: 0 forward sweep (EXTRAS: illumination / u.dt2 store), 2 adjoint sweep + imaging from stored u.dt2,
// 3 adjoint sweep + imaging by parts from the stored wavefield itself (B2FWI_HIST_UVDT2: grad -= u[t] * v.dt2[t]).
template <int R, int NPCC, int IMG, bool EXTRAS>
// (544 threads = 17 warps are allocated as 20 - warps come in groups of four - so ptxas budgets 96 registers; a
// direct __maxnreg__(120) compiles without the imaging variants' few spills but does not launch)
__global__ void __launch_bounds__(TmaCfg<R, NPCC, IMG>::NTHREADS, 1)
step3d_tma_kernel(const __grid_constant__ StepArgs a, const __grid_constant__ CUtensorMap m_cur,
                  const __grid_constant__ CUtensorMap m_prev, const __grid_constant__ CUtensorMap m_c1,
                  const __grid_constant__ CUtensorMap m_c2, const __grid_constant__ CUtensorMap m_h1)
{
    using C = TmaCfg<R, NPCC, IMG>;
    constexpr int TZ = C::TZ, TR = C::TR, ZH = C::ZH, SW = C::SW, NQ = C::NQ, NCUR = C::NCUR;
    extern __shared__ __align__(1024) unsigned char smem[];
    const uint32_t smem_s = (uint32_t)__cvta_generic_to_shared(smem);
    // layout: u[t] ring [NCUR][SROWS][SW] | operand ring [NPCC][3][TR][TZ] | mbarriers | weights | per-thread consts
    const uint32_t pcc_s = smem_s + NCUR * C::CUR_STRIDE;
    const uint32_t full_c = smem_s + C::BAR_OFF;                          // [NCUR]
    const uint32_t empty = full_c + NCUR * 8;                             // [NCUR]
    const uint32_t full_p = empty + NCUR * 8;                             // [NPCC]
    const uint32_t inj_full = full_p + NPCC * 8;                          // [NI] injection terms of a plane staged
    const uint32_t inj_empty = inj_full + C::NI * 8;                      // [NI] ... and consumed by every compute warp
    const uint32_t mask_bar = inj_empty + C::NI * 8;                      // plane mask written (one arrival per service warp)

    const int tid = threadIdx.x;
    const int ztile0 = blockIdx.x * TZ, r0 = blockIdx.y * TR;
    const int p_begin = (int)blockIdx.z * a.chunk;
    const int p_end = min(p_begin + a.chunk, a.np);
    const int n_it = p_end - p_begin;                    // planes computed by this CTA
    const int ncur = min(n_it + R, a.np - p_begin);      // u[t] planes staged (R more than computed: pipeline feed)

    // the tile lies inside the undamped interior (rows, z): planes blo_p..bhi_p-1 then need no c1
    const bool tile_in_box = __ldg(a.box + 6) == 1 && r0 >= __ldg(a.box + 2) && r0 + TR <= __ldg(a.box + 3) &&
                             ztile0 >= __ldg(a.box + 4) && ztile0 + TZ <= __ldg(a.box + 5);
    const int blo_p = tile_in_box ? __ldg(a.box + 0) : 0, bhi_p = tile_in_box ? __ldg(a.box + 1) : 0;

    // ---- fused sparse operators (tables in a.inj / a.itp; z range test only - no global loads on the CTA's start-up path)
    const int nrt = (a.nr + TR - 1) / TR;
    // (bounding box of the map's cells against this CTA's tile and chunk: a point source switches the injection path on
    // in the one to eight CTAs it touches, a receiver carpet in the z tile it lies in)
    const bool inj_on = a.inj.row_tile && a.inj.z_max >= ztile0 && a.inj.z_min < ztile0 + TZ &&
                        a.inj.r_max >= r0 && a.inj.r_min < r0 + TR && a.inj.p_max >= p_begin && a.inj.p_min < p_end;
    const bool itp_on = a.itp.row_tile && a.itp.z_max >= ztile0 && a.itp.z_min < ztile0 + TZ &&
                        a.itp.r_max >= r0 && a.itp.r_min < r0 + TR && a.itp.p_max >= p_begin && a.itp.p_min < p_end;
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < NCUR; s++) {
            mbar_init(full_c + 8 * s, 1);
            mbar_init(empty + 8 * s, C::NCONS / 32);
        }
#pragma unroll
        for (int s = 0; s < C::NI; s++) {
            mbar_init(inj_full + 8 * s, 1);
            mbar_init(inj_empty + 8 * s, C::NCONS / 32);
        }
        mbar_init(mask_bar, C::NSVC);
#pragma unroll
        for (int s = 0; s < NPCC; s++) mbar_init(full_p + 8 * s, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        float *wsm = reinterpret_cast<float *>(smem + C::W_OFF);
        wsm[0] = a.c0; wsm[1] = a.c0_lo;
#pragma unroll
        for (int k = 1; k <= R; k++) { wsm[1 + k] = a.cp[k]; wsm[1 + R + k] = a.cr[k]; wsm[1 + 2 * R + k] = a.cz[k]; }
    }
    if (tid < C::NCONS) {
        // Loop-invariant per-thread values take a round trip through shared memory: read back with a volatile
        // load they stay in registers, whereas ptxas re-derives them from %tid / the constant bank in every
        // plane otherwise (the sweep is issue-sensitive).
        const int tz = tid % C::TZQ, tr = tid / C::TZQ;
        uint4 c;
        c.x = smem_s + (uint32_t)((R + tr) * SW + ZH + 4 * tz) * 4u;       // this thread's float4 in u[t] stage 0
        c.y = pcc_s + (uint32_t)tid * 16u;                                 // ... in operand stage 0, array 0
        c.z = full_c;
        c.w = (uint32_t)(((int64_t)(r0 + tr) * a.sr + ztile0 + 4 * tz) >> 2);   // float4 index in plane 0 (halo 0)
        *reinterpret_cast<uint4 *>(smem + C::TH_OFF + tid * 16) = c;
    }
    __syncthreads();

    if (tid >= C::NCONS + 32) {
        // ------------------------------------------------------------------ service warps
        if (!inj_on && !itp_on) return;
        const int lane = tid & 31, sw = (tid - C::NCONS - 32) >> 5;
        const uint32_t imask_s = smem_s + C::IMASK_OFF, icnt_s = smem_s + C::ICNT_OFF, ient_s = smem_s + C::IENT_OFF;
        if (inj_on) {
            // Which planes of the chunk have injection cells in this tile's rows: one table lookup per plane while the
            // pipeline fills (words dealt round-robin to the three warps, two words per warp in flight: one memory
            // latency for a chunk of up to 192 planes). The compute warps test one bit per plane; only flagged planes go
            // through the staging handshake (a source touches two planes of a chunk, a receiver carpet every few).
            {
                const int rhi = min(r0 + TR, a.nr);
                for (int b = 32 * sw; b < n_it; b += 64 * C::NSVC) {
                    const int i0 = b + lane, i1 = b + 32 * C::NSVC + lane;
                    int a0 = 0, b0 = 0, a1 = 0, b1 = 0;
                    if (i0 < n_it) {
                        const int64_t row = (int64_t)(p_begin + i0) * a.nr;
                        a0 = __ldg(a.inj.con_rowptr + row + r0); b0 = __ldg(a.inj.con_rowptr + row + rhi);
                    }
                    if (i1 < n_it) {
                        const int64_t row = (int64_t)(p_begin + i1) * a.nr;
                        a1 = __ldg(a.inj.con_rowptr + row + r0); b1 = __ldg(a.inj.con_rowptr + row + rhi);
                    }
                    const uint32_t w0 = __ballot_sync(0xffffffffu, b0 > a0), w1 = __ballot_sync(0xffffffffu, b1 > a1);
                    if (lane == 0) {
                        asm volatile("st.shared.b32 [%0], %1;" ::"r"(imask_s + (uint32_t)(b >> 5) * 4u), "r"(w0) : "memory");
                        if (b + 32 * C::NSVC < n_it)
                            asm volatile("st.shared.b32 [%0], %1;" ::"r"(imask_s + (uint32_t)((b >> 5) + C::NSVC) * 4u), "r"(w1) : "memory");
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(mask_bar);
            }
            mbar_wait(mask_bar, 0);
            // k-th flagged plane -> warp k % NSVC, stage k % NI: one lane per row of the tile stages the row's terms in
            // ascending (cell, point) order for the compute thread that owns the cell
            int k = 0;
            for (int b = 0; b < n_it; b += 32) {
                uint32_t word;
                asm volatile("ld.shared.b32 %0, [%1];" : "=r"(word) : "r"(imask_s + (uint32_t)(b >> 5) * 4u));
                for (; word; word &= word - 1, k++) {
                    if (k % C::NSVC != sw) continue;
                    const int p = p_begin + b + __ffs(word) - 1, si = k % C::NI, row = r0 + lane;
                    const int64_t plane0 = (int64_t)p * a.sp;
                    int j0 = 0, j1 = 0;
                    if (lane < TR && row < a.nr) {
                        j0 = __ldg(a.inj.con_rowptr + (int64_t)p * a.nr + row);
                        j1 = __ldg(a.inj.con_rowptr + (int64_t)p * a.nr + row + 1);
                    }
                    if (k >= C::NI) mbar_wait(inj_empty + 8 * si, (uint32_t)(k / C::NI - 1) & 1u);    // stage consumed (flagged plane k - NI)
                    int cnt = 0;
                    for (int j = j0; j < j1; j++) {
                        const int64_t off = __ldg(a.inj.con_off + j);
                        const int z = (int)((uint32_t)(off - plane0) % (uint32_t)a.sr) - ztile0;
                        if (z >= 0 && z < TZ && cnt < C::NE) {
                            const float term = inject_term(__ldg(a.inj.contrib_w + j), __ldg(a.inj_vals + __ldg(a.inj.contrib_pt + j)),
                                                           a.dt, __ldg(a.vp + off));
                            asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(ient_s + (uint32_t)(((si * TR + lane) * C::NE + cnt) * 8)),
                                         "r"(__float_as_uint(term)), "r"((uint32_t)z) : "memory");
                            cnt++;
                        }
                    }
                    if (lane < TR) asm volatile("st.shared.b32 [%0], %1;" ::"r"(icnt_s + (uint32_t)((si * TR + lane) * 4)), "r"(cnt) : "memory");
                    __syncwarp();
                    if (lane == 0) mbar_arrive(inj_full + 8 * si);
                }
            }
        }
        // receivers whose home cell lies in this tile: rec[t][pt] = interpolate(u[t]). Reads `cur` from global memory
        // only (complete before the launch), so it needs no synchronisation with the sweep: lane <-> plane for the table
        // lookup, then the warp walks the (few) points of each plane that has any.
        if (itp_on) {
            for (int b = 32 * sw; b < n_it; b += 32 * C::NSVC) {
                const int i = b + lane;
                int lo = 0, hi = 0;
                if (i < n_it) {
                    const int key = (p_begin + i) * nrt + (int)blockIdx.y;
                    lo = __ldg(a.itp.pt_rowptr + key);
                    hi = __ldg(a.itp.pt_rowptr + key + 1);
                }
                for (uint32_t m = __ballot_sync(0xffffffffu, hi > lo); m; m &= m - 1) {
                    const int src_lane = __ffs(m) - 1;
                    const int plo = __shfl_sync(0xffffffffu, lo, src_lane), phi = __shfl_sync(0xffffffffu, hi, src_lane);
                    const int64_t plane0 = (int64_t)(p_begin + b + src_lane) * a.sp;
                    for (int q = plo + lane; q < phi; q += 32) {
                        const int z = (int)((uint32_t)(__ldg(a.itp.pt_home + q) - plane0) % (uint32_t)a.sr);
                        if (z >= ztile0 && z < ztile0 + TZ) {
                            const int pt = __ldg(a.itp.pt_order + q);
                            a.itp_out[pt] = interp_point(a.cur, pt, a.itp.ncorner, a.itp.corner_off, a.itp.corner_w);
                        }
                    }
                }
            }
        }
        return;
    }
    if (tid >= C::NCONS) {
        // ------------------------------------------------------------------ producer (one lane)
        if (tid == C::NCONS) {
            auto issue_cur = [&](int k) {
                const uint32_t bar = full_c + 8 * (k % NCUR);
                mbar_expect_tx(bar, C::CUR_BYTES);
                tma_load_3d(smem_s + (k % NCUR) * C::CUR_STRIDE, &m_cur, bar, ztile0 - ZH, r0 - R, p_begin + k);
            };
            auto issue_pcc = [&](int k) {
                const int p = p_begin + k;
                const bool skip_c1 = p >= blo_p && p < bhi_p;
                const uint32_t bar = full_p + 8 * (k % NPCC);
                const uint32_t dst = pcc_s + (k % NPCC) * C::PCC_BYTES;
                mbar_expect_tx(bar, (skip_c1 ? C::NARR - 1 : C::NARR) * C::PL_BYTES);
                tma_load_3d(dst, &m_prev, bar, ztile0, r0, p);
                tma_load_3d(dst + C::PL_BYTES, &m_c2, bar, ztile0, r0, p);
                if (!skip_c1) tma_load_3d(dst + 2 * C::PL_BYTES, &m_c1, bar, ztile0, r0, p);
                if (IMG >= 2) tma_load_3d(dst + 3 * C::PL_BYTES, &m_h1, bar, ztile0, r0, p);
            };
            // fill both rings, in the order the planes are needed
            for (int k = 0; k < NCUR; k++) {
                if (k < ncur) issue_cur(k);
                if (k < NPCC && k < n_it) issue_pcc(k);
            }
            for (int i = 0; i < n_it; i++) {             // iteration i has released stage i % NCUR / i % NPCC
                const bool need_cur = i + NCUR < ncur, need_pcc = i + NPCC < n_it;
                if (!need_cur && !need_pcc) break;
                mbar_wait(empty + 8 * (i % NCUR), (uint32_t)(i / NCUR) & 1u);
                if (need_pcc) issue_pcc(i + NPCC);
                if (need_cur) issue_cur(i + NCUR);
            }
        }
        return;
    }

    // ---------------------------------------------------------------------- consumers
    uint32_t ctr_s, pcc_t, bars, own4;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(ctr_s), "=r"(pcc_t), "=r"(bars), "=r"(own4) : "r"(smem_s + C::TH_OFF + tid * 16));
    StencilW<R> w;
    {
        const uint32_t ws = smem_s + C::W_OFF;
        asm volatile("ld.shared.f32 %0, [%1];" : "=f"(w.c0) : "r"(ws));
        asm volatile("ld.shared.f32 %0, [%1];" : "=f"(w.c0_lo) : "r"(ws + 4));
        w.cp[0] = w.cr[0] = w.cz[0] = 0.f;
#pragma unroll
        for (int k = 1; k <= R; k++) {
            asm volatile("ld.shared.f32 %0, [%1];" : "=f"(w.cp[k]) : "r"(ws + 4 * (1 + k)));
            asm volatile("ld.shared.f32 %0, [%1];" : "=f"(w.cr[k]) : "r"(ws + 4 * (1 + R + k)));
            asm volatile("ld.shared.f32 %0, [%1];" : "=f"(w.cz[k]) : "r"(ws + 4 * (1 + 2 * R + k)));
        }
    }
    const uint32_t bar_fc = bars, bar_em = bars + NCUR * 8, bar_fp = bars + 2 * NCUR * 8;
    const uint32_t bar_if = bar_fp + NPCC * 8, bar_ie = bar_if + C::NI * 8;
    uint32_t icur = 0, ipar = 0;     // injection stage of the next flagged plane as a byte offset into the row counts, and its parity
    if (inj_on) mbar_wait(bar_ie + C::NI * 8, 0);         // the plane mask is there (written by the service warps while the pipeline filled)
    const int tz = tid % C::TZQ, tr = tid / C::TZQ;
    const bool active = (r0 + tr < a.nr) && (ztile0 + tz * 4 < a.nz);
    const bool lane0 = (tid & 31) == 0;
    const uint32_t sp4 = (uint32_t)(a.sp >> 2);
    const float4 one4 = make_float4(1.f, 1.f, 1.f, 1.f);
#define F4(ptr) reinterpret_cast<const float4 *>(ptr)
#define F4W(ptr) reinterpret_cast<float4 *>(ptr)

    // register pipeline: plane offset o in -R..R of iteration (unroll slot) j lives in q[(j + R + o) % NQ]
    float4 q[NQ];
#pragma unroll
    for (int i = 0; i < R; i++) {
        const int p = p_begin - R + i;
        q[i] = (active && p >= 0) ? F4(a.cur)[own4 + (uint32_t)p * sp4] : zero4();
    }
#pragma unroll
    for (int i = 0; i < R; i++) {
        if (i < ncur) {
            mbar_wait(bar_fc + 8 * i, 0);
            q[R + i] = lds4(ctr_s + i * C::CUR_STRIDE);
        } else {
            q[R + i] = zero4();
        }
    }

    uint32_t idx = own4 + (uint32_t)p_begin * sp4;
    uint32_t ps = 0, pp = 0;         // operand-ring stage / parity when they are not compile-time (see TmaCfg::PCT)
    // one group = NQ planes. EDGE = false: the whole group and its feed planes exist, no per-plane range tests.
    auto group = [&](auto edge_tag, int pb, uint32_t par) {
        constexpr bool EDGE = decltype(edge_tag)::value;
#pragma unroll
        for (int j = 0; j < NQ; j++) {
            const int i = pb + j;
            if (!EDGE || i < n_it) {
                float4 g4;
                // imaging: grad goes through registers (the operand ring has no room for a fifth array); the load
                // is issued before the barrier waits and consumed after the update
                if (IMG >= 2 && active) g4 = F4(a.grad)[idx];
                // feed: plane i+R enters the register pipeline from its ring stage
                const int sf = (j + R) % NQ;
                if (!EDGE || i + R < ncur) {
                    mbar_wait(bar_fc + 8 * sf, ((j + R) / NQ) ? par ^ 1u : par);
                    q[(j + 2 * R) % NQ] = lds4(ctr_s + sf * C::CUR_STRIDE);
                } else {
                    q[(j + 2 * R) % NQ] = zero4();
                }
                // pointwise operands of plane i
                uint32_t pc;
                if (C::PCT) {
                    mbar_wait(bar_fp + 8 * (j % NPCC), par ^ (uint32_t)((j / NPCC) & 1));
                    pc = pcc_t + (j % NPCC) * C::PCC_BYTES;
                } else {
                    mbar_wait(bar_fp + 8 * ps, pp);
                    pc = pcc_t + ps * C::PCC_BYTES;
                    if (++ps == NPCC) { ps = 0; pp ^= 1u; }
                }
                const int p = p_begin + i;
                float4 prev, c2, c1, h1, o;
                if (IMG == 3) {
                    // the by-parts imaging keeps u[t-1] alive to the end: fetch the pointwise operands after the
                    // Laplacian instead of holding 16 registers across it (the variant is register-bound at 96)
                    const float4 lap = point_laplacian<R, 3, SW>(w, q, j, SAddr{ctr_s + (j % NCUR) * C::CUR_STRIDE});
                    prev = lds4(pc);
                    c2 = lds4(pc + C::PL_BYTES);
                    c1 = (p >= blo_p && p < bhi_p) ? one4 : lds4(pc + 2 * C::PL_BYTES);
                    h1 = lds4(pc + 3 * C::PL_BYTES);
                    o = point_finish<false>(q[(j + R) % NQ], lap, prev, c1, c2, 4);
                } else {
                    prev = lds4(pc);
                    c2 = lds4(pc + C::PL_BYTES);
                    c1 = (p >= blo_p && p < bhi_p) ? one4 : lds4(pc + 2 * C::PL_BYTES);
                    if (IMG == 2) h1 = lds4(pc + 3 * C::PL_BYTES);
                    // no z masking: beyond nz the TMA unit filled u[t], u[t-1] and c2 with zeros, so o == 0 there
                    o = point_update<R, 3, SW, false>(w, q, j, SAddr{ctr_s + (j % NCUR) * C::CUR_STRIDE}, prev, c1, c2, 4);
                }
                bool inj_here = false;
                if (inj_on) {        // CTA-uniform; one mask bit per plane
                    uint32_t word;
                    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(word) : "r"(smem_s + C::IMASK_OFF + (uint32_t)(i >> 5) * 4u));
                    inj_here = (word >> (i & 31)) & 1u;
                }
                if (inj_here) {
                    // injection terms of this plane, staged by a service warp (a row rarely has any)
                    mbar_wait(bar_if + (icur >> 3), ipar);
                    uint32_t n;
                    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(n) : "r"(smem_s + C::ICNT_OFF + icur + 4u * (uint32_t)tr));
                    for (uint32_t e = 0; e < n; e++) {
                        uint32_t tb, zl;
                        asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];" : "=r"(tb), "=r"(zl)
                                     : "r"(smem_s + C::IENT_OFF + (icur * 2u + 8u * (uint32_t)tr) * C::NE + 8u * e));
                        if ((int)(zl >> 2) == tz) {
                            const float t = __uint_as_float(tb);
                            if ((zl & 3u) == 0u) o.x = __fadd_rn(o.x, t);
                            else if ((zl & 3u) == 1u) o.y = __fadd_rn(o.y, t);
                            else if ((zl & 3u) == 2u) o.z = __fadd_rn(o.z, t);
                            else o.w = __fadd_rn(o.w, t);
                        }
                    }
                }
                // stage j (u[t] plane i), the operand stage and the injection stage are free again
                __syncwarp();
                if (lane0) {
                    mbar_arrive(bar_em + 8 * (j % NCUR));
                    if (inj_here) mbar_arrive(bar_ie + (icur >> 3));
                }
                if (inj_here) {
                    icur += 4u * TR;
                    if (icur == 4u * TR * C::NI) { icur = 0; ipar ^= 1u; }
                }
                if (active) {
                    F4W(a.out)[idx] = o;
                    if (IMG == 2) F4W(a.grad)[idx] = img4(g4, h1, q[(j + R) % NQ]);
                    if (IMG == 3) F4W(a.grad)[idx] = img4(g4, h1, d2u4(prev, q[(j + R) % NQ], o, a.inv_dt2));
                    if (EXTRAS) {
                        const float4 Cc = q[(j + R) % NQ];
                        if (a.illum) F4W(a.illum)[idx] = fma4(Cc, Cc, F4(a.illum)[idx]);
                        if (a.d2u) F4W(a.d2u)[idx] = d2u4(prev, Cc, o, a.inv_dt2);
                    }
                }
                idx += sp4;
            }
        }
    };
    uint32_t par = 0;
    for (int pb = 0; pb < n_it; pb += NQ, par ^= 1u) {
        if (pb + NQ <= n_it && pb + NQ + R <= ncur) group(std::false_type{}, pb, par);
        else group(std::true_type{}, pb, par);
    }
#undef F4
#undef F4W
}

// ---- host side: tensor maps through the driver entry point (no link-time dependency on libcuda)
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn()
{
    static EncodeTiledFn fn = []() -> EncodeTiledFn {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            return nullptr;
        return reinterpret_cast<EncodeTiledFn>(p);
    }();
    return fn;
}

static int make_map(CUtensorMap *m, const float *field, const StepArgs &a, int box_z, int box_r)
{
    EncodeTiledFn enc = encode_fn();
    if (!enc) { set_error("cuTensorMapEncodeTiled is not available from the driver"); return B2FWI_ECUDA; }
    const cuuint64_t dims[3] = {(cuuint64_t)a.nz, (cuuint64_t)a.nr, (cuuint64_t)a.np};
    const cuuint64_t strides[2] = {(cuuint64_t)a.sr * 4, (cuuint64_t)a.sp * 4};
    const cuuint32_t box[3] = {(cuuint32_t)box_z, (cuuint32_t)box_r, 1};
    const cuuint32_t es[3] = {1, 1, 1};
    const CUresult rc = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float *>(field), dims, strides, box, es,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (rc != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed (%d)", (int)rc); return B2FWI_ECUDA; }
    return 0;
}

template <int R, int NPCC, int IMG, bool EXTRAS>
static int launch_tma(const StepArgs &a, cudaStream_t st)
{
    using C = TmaCfg<R, NPCC, IMG>;
    static_assert(C::SMEM <= 232448, "shared memory per CTA");
    auto kern = step3d_tma_kernel<R, NPCC, IMG, EXTRAS>;
    static bool configured = false;
    if (!configured) {
        B2_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM));
        configured = true;
    }
    CUtensorMap m_cur, m_prev, m_c1, m_c2, m_h1;
    int rc;
    if ((rc = make_map(&m_cur, a.cur, a, C::SW, C::SROWS))) return rc;
    if ((rc = make_map(&m_prev, a.prev, a, C::TZ, C::TR))) return rc;
    if ((rc = make_map(&m_c1, a.c1, a, C::TZ, C::TR))) return rc;
    if ((rc = make_map(&m_c2, a.c2, a, C::TZ, C::TR))) return rc;
    if ((rc = make_map(&m_h1, IMG >= 2 ? a.h1 : a.c2, a, C::TZ, C::TR))) return rc;
    const int nchunks = (a.np + a.chunk - 1) / a.chunk;
    dim3 grid((a.nz + C::TZ - 1) / C::TZ, (a.nr + C::TR - 1) / C::TR, nchunks);
    kern<<<grid, C::NTHREADS, C::SMEM, st>>>(a, m_cur, m_prev, m_c1, m_c2, m_h1);
    B2_CUDA(cudaGetLastError());
    count_launch();
    return 0;
}

// B2FWI_TMA: bit 0 forward sweep, bit 1 adjoint + imaging sweep (A/B against the register-staged kernels)
static int g_tma = []() { const char *e = getenv("B2FWI_TMA"); return e ? atoi(e) : 3; }();
void set_tma_mask(int mask) { g_tma = mask & 3; }
int get_tma_mask() { return g_tma; }

bool tma_step_supported(const Layout &L, const StepArgs &a, int img)
{
    if (L.ndim != 3 || !(img == 0 || img == 2) || L.halo != 0 || L.R < 1 || L.R > 8) return false;
    if (L.fs) return false;      // free-surface models run on the register-staged kernels (mirrored top rows, stream_point.cuh)
    if (!(g_tma & (img == 0 ? 1 : 2))) return false;
    if (img == 2 && (!a.h1 || !a.grad || (((uintptr_t)a.h1 | (uintptr_t)a.grad) & 15))) return false;
    if (img == 2 && (a.illum || a.d2u)) return false;
    // TMA: 16-byte aligned base and strides (rows are pitched to 32 floats); float4 stores as in step_kernel
    const uintptr_t al = (uintptr_t)a.cur | (uintptr_t)a.prev | (uintptr_t)a.c1 | (uintptr_t)a.c2 | (uintptr_t)a.out;
    return (al & 15) == 0 && (L.sr % 4) == 0 && encode_fn() != nullptr;
}

// B2FWI_FUSE / b2fwi_set_option("fuse", 0|1): sparse operators inside the sweep kernels (default) or as separate launches
// bit 0: injection of small maps (sources: <= FUSE_SMALL cells), bit 1: interpolation, bit 2: injection of any map.
// Default 7 (everything inside the sweep), from interleaved runs on one box (592^3 so=8, 688 steps, 16384 receivers):
// forward sweep with source injection + receiver interpolation 0.422 s fused against 0.428 s as three launches per
// step; recompute + adjoint with the residual injection of the receiver carpet 1.057 s either way (the tiles of the
// receivers' z range test the plane mask and go through the staging handshake, which costs about what the 6 us
// kernel and its launch gap did). Only the CTAs inside the bounding box of a map's cells take part at all.
static int g_fuse = []() { const char *e = getenv("B2FWI_FUSE"); return e ? (atoi(e) & 7) : 7; }();
static const int FUSE_SMALL = 64;
void set_fuse(int mask) { g_fuse = mask & 7; }
int get_fuse() { return g_fuse; }
bool tma_fusable(const Layout &L, const b2fwi_sparse *m, int what)
{
    const bool want = (what == 2) ? (g_fuse & 2) : ((g_fuse & 4) || ((g_fuse & 1) && m && m->ncell <= FUSE_SMALL));
    return want && m && m->row_tile == 16 && m->con_rowptr && m->con_off && m->pt_order && m->pt_home &&
           m->pt_rowptr && m->max_row_con <= (L.R <= 4 ? 8 : 4) && L.ndim == 3 && L.halo == 0;
}

// tile of the TMA kernels (pick_chunk sizes the plane chunks for whole waves of one CTA per SM)
void tma_tile_shape(int R, int *tz, int *tr)
{
    *tz = (R <= 4) ? 128 : 64;
    *tr = 16;
}

bool tma_enabled(int img) { return (g_tma & (img == 0 ? 1 : 2)) != 0 && encode_fn() != nullptr; }

template <int R, int NPCC>
static int launch_tma_r(const StepArgs &a, int img, cudaStream_t st)
{
    // (so = 4: the 5-stage operand ring of the forward sweep does not fit next to a fourth operand array)
    constexpr int NPI = (R == 2) ? 3 : NPCC;
    if (img == 2) return a.hist_uv ? launch_tma<R, NPI, 3, false>(a, st) : launch_tma<R, NPI, 2, false>(a, st);
    return (a.illum || a.d2u) ? launch_tma<R, NPCC, 0, true>(a, st) : launch_tma<R, NPCC, 0, false>(a, st);
}

int launch_step_tma(const Layout &L, const StepArgs &a, int img, cudaStream_t st)
{
    switch (L.R) {
    case 1: return launch_tma_r<1, 3>(a, img, st);
    case 2: return launch_tma_r<2, 5>(a, img, st);
    case 3: return launch_tma_r<3, 3>(a, img, st);
    case 4: return launch_tma_r<4, 3>(a, img, st);
    case 5: return launch_tma_r<5, 3>(a, img, st);
    case 6: return launch_tma_r<6, 3>(a, img, st);
    case 7: return launch_tma_r<7, 3>(a, img, st);
    case 8: return launch_tma_r<8, 3>(a, img, st);
    default: set_error("no TMA variant for stencil radius %d", L.R); return B2FWI_EUNSUPPORTED;
    }
}

}  // namespace b2fwi

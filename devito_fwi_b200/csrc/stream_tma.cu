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
    static constexpr int NCONS = TZQ * TR, NTHREADS = NCONS + 32;
    static constexpr int BAR_OFF = NCUR * CUR_STRIDE + NPCC * PCC_BYTES;    // mbarriers (64 slots reserved)
    static constexpr int W_OFF = BAR_OFF + 64 * 8;                          // Laplacian weights, 32 floats
    static constexpr int TH_OFF = W_OFF + 32 * 4;                           // per-thread constants, uint4 each
    static constexpr int SMEM = TH_OFF + NCONS * 16;
    static_assert(R >= 1 && R <= 8, "stencil radius");
    static_assert(PL_BYTES % 128 == 0, "TMA destinations are 128-byte aligned");
    static_assert(2 * NCUR + NPCC <= 64 && 2 + 3 * R <= 32, "barrier / weight slots");
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

    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < NCUR; s++) {
            mbar_init(full_c + 8 * s, 1);
            mbar_init(empty + 8 * s, C::NCONS / 32);
        }
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
                // stage j (u[t] plane i) and the operand stage are free again
                __syncwarp();
                if (lane0) mbar_arrive(bar_em + 8 * (j % NCUR));
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
    if (!(g_tma & (img == 0 ? 1 : 2))) return false;
    if (img == 2 && (!a.h1 || !a.grad || (((uintptr_t)a.h1 | (uintptr_t)a.grad) & 15))) return false;
    if (img == 2 && (a.illum || a.d2u)) return false;
    // TMA: 16-byte aligned base and strides (rows are pitched to 32 floats); float4 stores as in step_kernel
    const uintptr_t al = (uintptr_t)a.cur | (uintptr_t)a.prev | (uintptr_t)a.c1 | (uintptr_t)a.c2 | (uintptr_t)a.out;
    return (al & 15) == 0 && (L.sr % 4) == 0 && encode_fn() != nullptr;
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
    if (img == 2) return a.hist_uv ? launch_tma<R, NPCC, 3, false>(a, st) : launch_tma<R, NPCC, 2, false>(a, st);
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

// qw2d.cu -- quadratic-Wasserstein misfit of whole shot records on the device: the back-and-forth method of the
// reference's QW2D solver (misfit/QW2D/src/fot2d.c, driven per shot through files and a subprocess by
// misfit/bfm.py:145-193 and misfit/misfit.py:69-104), for all shots of a rank at once.
//
// What the reference computes per record f = syn, g = obs of shape (nt, nrec), n1 = nrec contiguous, n2 = nt:
//   misfit.py:18-45   shift positive:  c = gamma * max(0, -min(f, g));  f += c, g += c;  mass = mean(f)
//   fot2d.c:608-656   mu = f / mean(f), nu = g / mean(g);  sigma = step_scale / max(mu, nu)
//   fot2d.c:514-606   num_steps times:  phi  += sigma * (-Laplace)^-1 (rho - nu)   (DCT Poisson solve, :479-503)
//                                       (phi, dual) <- c-transforms (row / column Legendre transforms over lower
//                                       convex hulls, :66-178);  sigma update (:505-517);
//                                       rho <- push-forward of nu by grad phi (:290-478);   then the same with the
//                                       roles (dual, mu);   W2 value (:519-531) after each half
//   fot2d.c:636-655   adjoint source = (dual - <mu, dual>) / mean(f)  and  misfit.py:79: / mass
// The solver is single precision with double-precision sub-expressions wherever C's usual arithmetic conversions
// put them; this file follows those conversions expression by expression (the file is compiled with -fmad=false:
// the reference is built with -std=c11, i.e. without contraction), and it accumulates the sums that steer the step
// size (W2 value, H^-1 residual, mass of rho) in the reference's order - sequentially, in float - because the
// sigma update compares differences of those sums that are close to their rounding noise. Two things are NOT
// bit-reproducible by construction: the DCTs (FFTW there, cuFFT here; both accurate to float rounding) and the
// accumulation order inside the push-forward (the reference's own order depends on its OpenMP schedule; here a
// 64-bit fixed-point accumulator makes the result independent of the order).
#include <cufft.h>
#include <math.h>
#include <string.h>

#include <map>
#include <mutex>
#include <tuple>

#include "common.cuh"

namespace b2fwi {
namespace qw {

struct Scal {            // per-record scalars (device)
    float c;             // positivity shift
    float mass;          // mean of the shifted synthetic record, numpy float32 pairwise sum / size (misfit.py:73)
    float sum1, sum2;    // means of f and g (normalize.c:14-27)
    float sigma, value, old_value, grad_sq;
    float rho_sum, term, loss;
    int skip;            // sum1 <= 0: fotGradient2d returns 0 and leaves the adjoint source at 1 (fot2d.c:626-627)
};

struct Ptrs {            // per-call device pointers; record s at offset s * pc (s * mc for the maps)
    float *mu, *nu, *phi, *dual, *rho, *ws, *tmp, *xmap, *ymap, *kern;
    int2 *hull;          // hull stacks of the c-transform lines: (point index, value)
    long long *racc;
    float *freal;        // cuFFT real buffer  [lines][N]
    float2 *fcplx;       // cuFFT complex buffer [lines][N/2+1]
    Scal *sc;
};

#define QW_FIXED_SCALE 1099511627776.0f          /* 2^40 */
#define QW_FIXED_INV 9.094947017729282e-13      /* 2^-40 */

// ------------------------------------------------------------------------------------------------ set-up
// min over a record of (syn - dw) and (obs - dw)   (misfit.py:20: min(f.min(), g.min()))
__global__ void qw_min_kernel(const float *__restrict__ syn, const float *__restrict__ obs, const float *__restrict__ dw,
                              int64_t pc, double gamma, Scal *__restrict__ sc)
{
    __shared__ float sh[256];
    const int s = blockIdx.x;
    const int64_t base = (int64_t)s * pc;
    float m = 3.4e38f;
    for (int64_t i = threadIdx.x; i < pc; i += blockDim.x) {
        const float d = dw ? dw[base + i] : 0.f;
        m = fminf(m, fminf(syn[base + i] - d, obs[base + i] - d));
    }
    sh[threadIdx.x] = m;
    __syncthreads();
    for (int k = 128; k > 0; k >>= 1) {
        if (threadIdx.x < k) sh[threadIdx.x] = fminf(sh[threadIdx.x], sh[threadIdx.x + k]);
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const float mn = sh[0];
        // c = (-min) * gamma in fp32: numpy's result for an fp32 scalar times a python float (NEP 50), as in w1d
        sc[s].c = (mn < 0.f) ? __fmul_rn(-mn, (float)gamma) : 0.f;
    }
}

// f = (syn - dw) + c, g = (obs - dw) + c   (fwi.py:146-150, misfit.py:41-42), stored in mu / nu
__global__ void qw_shift_kernel(const float *__restrict__ syn, const float *__restrict__ obs, const float *__restrict__ dw,
                                int64_t pc, int ns, const Scal *__restrict__ sc, float *__restrict__ f, float *__restrict__ g)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= pc * ns) return;
    const float c = sc[i / pc].c;
    const float d = dw ? dw[i] : 0.f;
    f[i] = __fadd_rn(__fsub_rn(syn[i], d), c);
    g[i] = __fadd_rn(__fsub_rn(obs[i], d), c);
}

// numpy float32 sum() of a contiguous array: pairwise summation, 8 accumulators per <= 128-element block
// (numpy/core/src/umath/loops_utils.h.src). The recursion of the original is unrolled on an explicit stack: a
// 510 k-element record nests 12 levels deep, more than the default device stack takes.
static __device__ float np_block_sum(const float *v, int n)
{
    if (n < 8) {
        float res = 0.f;
        for (int i = 0; i < n; i++) res = __fadd_rn(res, v[i]);
        return res;
    }
    float r[8];
    for (int j = 0; j < 8; j++) r[j] = v[j];
    int i;
    for (i = 8; i < n - (n % 8); i += 8)
        for (int j = 0; j < 8; j++) r[j] = __fadd_rn(r[j], v[i + j]);
    float res = __fadd_rn(__fadd_rn(__fadd_rn(r[0], r[1]), __fadd_rn(r[2], r[3])),
                          __fadd_rn(__fadd_rn(r[4], r[5]), __fadd_rn(r[6], r[7])));
    for (; i < n; i++) res = __fadd_rn(res, v[i]);
    return res;
}

static __device__ float np_pairwise_sum(const float *v, int64_t n)
{
    int64_t off[48], len[48];
    float left[48];
    int stage[48];
    int sp = 0;
    off[0] = 0; len[0] = n; stage[0] = 0;
    float ret = 0.f;
    while (sp >= 0) {
        if (len[sp] <= 128) { ret = np_block_sum(v + off[sp], (int)len[sp]); sp--; continue; }
        int64_t n2 = len[sp] / 2;
        n2 -= n2 % 8;
        if (stage[sp] == 0) {                     // descend into the left half
            stage[sp] = 1;
            off[sp + 1] = off[sp]; len[sp + 1] = n2; stage[sp + 1] = 0;
            sp++;
        } else if (stage[sp] == 1) {              // left half done: keep it, descend into the right half
            left[sp] = ret;
            stage[sp] = 2;
            off[sp + 1] = off[sp] + n2; len[sp + 1] = len[sp] - n2; stage[sp + 1] = 0;
            sp++;
        } else {                                  // both halves done
            ret = __fadd_rn(left[sp], ret);
            sp--;
        }
    }
    return ret;
}

// f.sum() / f.size (misfit.py:73) with numpy's pairwise summation tree, one block per record: the tree splits at
// n/2 rounded down to a multiple of 8 until a node has <= 128 elements, so its top MASS_DEPTH levels are walked by
// 2^MASS_DEPTH threads in parallel - each sums its own subtree exactly as numpy would - and the partial sums are
// combined pairwise in the tree's own order (left + right at every split node).
#define MASS_DEPTH 8
#define QW_PF 8              // read-ahead of the hull walk, points
__global__ void __launch_bounds__(1 << MASS_DEPTH) qw_mass_kernel(const float *__restrict__ f, int64_t pc, int ns,
                                                                  Scal *__restrict__ sc)
{
    __shared__ float val[1 << MASS_DEPTH];
    const int s = blockIdx.x, t = threadIdx.x;
    const float *v = f + (int64_t)s * pc;
    int64_t off = 0, len = pc;
    int leaf_level = MASS_DEPTH;              // level at which this thread's path reaches an unsplit node
    for (int level = 0; level < MASS_DEPTH; level++) {
        if (len <= 128) { leaf_level = level; break; }
        int64_t n2 = len / 2;
        n2 -= n2 % 8;
        if ((t >> (MASS_DEPTH - 1 - level)) & 1) { off += n2; len -= n2; }
        else len = n2;
    }
    // an unsplit node above the bottom level is shared by 2^(MASS_DEPTH - leaf_level) threads: the first one sums it
    const bool owner = (t & ((1 << (MASS_DEPTH - leaf_level)) - 1)) == 0;
    val[t] = owner ? np_pairwise_sum(v + off, len) : 0.f;
    __syncthreads();
    for (int level = MASS_DEPTH - 1; level >= 0; level--) {
        const int span = 1 << (MASS_DEPTH - level);
        if ((t & (span - 1)) == 0 && level < leaf_level) val[t] = __fadd_rn(val[t], val[t + span / 2]);
        __syncthreads();
    }
    if (t == 0) sc[s].mass = __fdiv_rn(val[0], (float)pc);
}

// ------------------------------------------------------------------------------------------------ sequential sums
// s = fl(s + t_i), i ascending: the reference's `for (i...) sum += expr;` loops in float. One block per record; the
// block computes the terms of a chunk in parallel into shared memory, thread 0 adds them up in order while the other
// threads already compute the next chunk.
enum { SEQ_SUM_F = 0, SEQ_SUM_G, SEQ_W2, SEQ_H1, SEQ_RHO, SEQ_TERM };
#define SEQ_CHUNK 2048

template <int WHAT>
static __device__ __forceinline__ float seq_term(const Ptrs &p, int64_t base, int i, int n1, int n2, int pc,
                                                 const float *__restrict__ other)
{
    if (WHAT == SEQ_SUM_F) return p.mu[base + i];                                   // normalize.c:18 sum1 += f[i]
    if (WHAT == SEQ_SUM_G) return p.nu[base + i];                                   // normalize.c:19
    if (WHAT == SEQ_RHO) return __fdiv_rn(p.rho[base + i], (float)pc);              // fot2d.c:470 sum += rho[i]/pcount
    if (WHAT == SEQ_TERM) return __fdiv_rn(__fmul_rn(p.mu[base + i], p.dual[base + i]), (float)pc);   // fot2d.c:643
    if (WHAT == SEQ_H1)                                                              // fot2d.c:498 h1 += ws*(rho-nu)
        return __fmul_rn(p.ws[base + i], __fsub_rn(p.rho[base + i], other[base + i]));
    // SEQ_W2, fot2d.c:526-527: .5*(x*x+y*y)*(mu+nu) - nu*phi - mu*dual with x, y float, .5 double
    const int r = i / n1, j = i - r * n1;
    const float x = (float)((j + .5) / (n1 * 1.0)), y = (float)((r + .5) / (n2 * 1.0));
    const float mu = p.mu[base + i], nu = p.nu[base + i];
    const double a = .5 * (double)__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)) * (double)__fadd_rn(mu, nu);
    const double t = a - (double)__fmul_rn(nu, p.phi[base + i]) - (double)__fmul_rn(mu, p.dual[base + i]);
    // the reference adds this double to its float accumulator in double and rounds ((float)((double)v + t)); adding the
    // float-rounded term differs from that only when v + t falls within 2^-25 |t| of a rounding boundary
    return (float)t;
}

template <int WHAT>
__global__ void __launch_bounds__(512) qw_seq_kernel(Ptrs p, int n1, int n2, const float *__restrict__ other)
{
    __shared__ __align__(16) float buf[2][SEQ_CHUNK];
    const int s = blockIdx.x, pc = n1 * n2;
    const int64_t base = (int64_t)s * pc;
    float acc = 0.f;
    const int nchunk = (pc + SEQ_CHUNK - 1) / SEQ_CHUNK;
    for (int i = threadIdx.x; i < SEQ_CHUNK && i < pc; i += blockDim.x) buf[0][i] = seq_term<WHAT>(p, base, i, n1, n2, pc, other);
    __syncthreads();
    // Warp 0 does nothing but the sequential sum (lane 0; its other lanes idle - in a shared warp the two divergent
    // paths would take turns at the issue slot and the dependent-add chain, the floor of this kernel at 4 cycles per
    // element, would wait for the staging code); warps 1.. stage the next chunk meanwhile.
    const int nstage = (int)blockDim.x - 32;
    for (int c = 0; c < nchunk; c++) {
        const int cnt = min(SEQ_CHUNK, pc - c * SEQ_CHUNK);
        if (threadIdx.x == 0) {
            const float *b = buf[c & 1];
            const float4 *b4 = reinterpret_cast<const float4 *>(b);
            int i = 0;
            for (; i + 16 <= cnt; i += 16) {
                const float4 v0 = b4[(i >> 2) + 0], v1 = b4[(i >> 2) + 1], v2 = b4[(i >> 2) + 2], v3 = b4[(i >> 2) + 3];
                acc = __fadd_rn(acc, v0.x); acc = __fadd_rn(acc, v0.y); acc = __fadd_rn(acc, v0.z); acc = __fadd_rn(acc, v0.w);
                acc = __fadd_rn(acc, v1.x); acc = __fadd_rn(acc, v1.y); acc = __fadd_rn(acc, v1.z); acc = __fadd_rn(acc, v1.w);
                acc = __fadd_rn(acc, v2.x); acc = __fadd_rn(acc, v2.y); acc = __fadd_rn(acc, v2.z); acc = __fadd_rn(acc, v2.w);
                acc = __fadd_rn(acc, v3.x); acc = __fadd_rn(acc, v3.y); acc = __fadd_rn(acc, v3.z); acc = __fadd_rn(acc, v3.w);
            }
            for (; i < cnt; i++) acc = __fadd_rn(acc, b[i]);
        } else if (threadIdx.x >= 32 && c + 1 < nchunk) {
            const int i0 = (c + 1) * SEQ_CHUNK;
            for (int i = threadIdx.x - 32; i < SEQ_CHUNK && i0 + i < pc; i += nstage)
                buf[(c + 1) & 1][i] = seq_term<WHAT>(p, base, i0 + i, n1, n2, pc, other);
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        Scal &q = p.sc[s];
        if (WHAT == SEQ_SUM_F) q.sum1 = __fdiv_rn(acc, (float)pc);                  // normalize.c:22 (flag = 1)
        if (WHAT == SEQ_SUM_G) q.sum2 = __fdiv_rn(acc, (float)pc);
        if (WHAT == SEQ_W2) q.value = __fdiv_rn(acc, (float)pc);                    // fot2d.c:530 value /= pcount
        if (WHAT == SEQ_H1) q.grad_sq = __fdiv_rn(acc, (float)pc);                  // fot2d.c:501 h1 /= pcount
        if (WHAT == SEQ_RHO) q.rho_sum = acc;
        if (WHAT == SEQ_TERM) q.term = acc;
    }
}

// ------------------------------------------------------------------------------------------------ elementwise
// fot2d.c:626-627,629-632 + fot2d.c:222-236: mu, nu normalised; phi = dual = .5 (x^2 + y^2); rho = mu; max(mu, nu)
__global__ void qw_normalize_kernel(Ptrs p, int n1, int n2, int ns)
{
    const int pc = n1 * n2;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)pc * ns) return;
    const int s = (int)(i / pc), k = (int)(i - (int64_t)s * pc);
    const Scal &q = p.sc[s];
    const float m = (q.sum1 > 0.f) ? __fdiv_rn(p.mu[i], q.sum1) : 0.f;
    const float n = (q.sum2 > 0.f) ? __fdiv_rn(p.nu[i], q.sum2) : 0.f;
    p.mu[i] = m;
    p.nu[i] = n;
    p.rho[i] = m;
    const int r = k / n1, j = k - r * n1;
    const float x = (float)((j + .5) / (n1 * 1.0)), y = (float)((r + .5) / (n2 * 1.0));
    const float z = (float)(.5 * (double)__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)));
    p.phi[i] = z;
    p.dual[i] = z;
}

// sigma = step_scale / fmax(max mu, max nu)   (fot2d.c:238-248,634-636); also flags empty records
__global__ void qw_sigma_kernel(Ptrs p, int pc, float step_scale)
{
    __shared__ float sh[256];
    const int s = blockIdx.x;
    const int64_t base = (int64_t)s * pc;
    float m = 0.f;
    for (int i = threadIdx.x; i < pc; i += blockDim.x) m = fmaxf(m, fmaxf(p.mu[base + i], p.nu[base + i]));
    sh[threadIdx.x] = m;
    __syncthreads();
    for (int k = 128; k > 0; k >>= 1) {
        if (threadIdx.x < k) sh[threadIdx.x] = fmaxf(sh[threadIdx.x], sh[threadIdx.x + k]);
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        Scal &q = p.sc[s];
        q.skip = !(q.sum1 > 0.f);
        q.sigma = __fdiv_rn(step_scale, sh[0]);
    }
}

// negative-Laplacian symbol (fot2d.c:4-18)
__global__ void qw_kernel_kernel(float *__restrict__ kern, int n1, int n2)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n1 * n2) return;
    const int r = i / n1, j = i - r * n1;
    const float x = (float)(M_PI * j / (n1 * 1.0)), y = (float)(M_PI * r / (n2 * 1.0));
    kern[i] = (float)((double)(2 * n1 * n1) * (1 - cos((double)x)) + (double)(2 * n2 * n2) * (1 - cos((double)y)));
}

// ws = rho - other   (fot2d.c:482-484)
__global__ void qw_sub_kernel(const float *__restrict__ rho, const float *__restrict__ other, float *__restrict__ ws, int64_t n)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) ws[i] = __fsub_rn(rho[i], other[i]);
}

// ws[0] = 0; ws[i] /= 4 * pcount * kernel[i]   (fot2d.c:488-491)
__global__ void qw_poisson_scale_kernel(float *__restrict__ ws, const float *__restrict__ kern, int pc, int ns)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)pc * ns) return;
    const int k = (int)(i % pc);
    ws[i] = (k == 0) ? 0.f : __fdiv_rn(ws[i], __fmul_rn((float)(4 * pc), kern[k]));
}

// pot += sigma * ws   (fot2d.c:496-497)
__global__ void qw_axpy_kernel(float *__restrict__ pot, const float *__restrict__ ws, const Scal *__restrict__ sc, int pc, int ns)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)pc * ns) return;
    const Scal &q = sc[i / pc];
    if (q.skip) return;
    pot[i] = __fadd_rn(pot[i], __fmul_rn(q.sigma, ws[i]));
}

// sigma update and bookkeeping after a half step (fot2d.c:505-517,553-555,593-595)
__global__ void qw_step_update_kernel(Scal *__restrict__ sc, int ns)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= ns) return;
    Scal &q = sc[s];
    const float scale_down = .8f;                        // float scaleDown = .8
    const float scale_up = (float)(1. / (double)scale_down);   // scaleUp = 1./scaleDown
    const float upper = .75f, lower = .25f;
    const float diff = __fsub_rn(q.value, q.old_value);
    if (diff > __fmul_rn(__fmul_rn(q.grad_sq, q.sigma), upper)) q.sigma = __fmul_rn(q.sigma, scale_up);
    else if (diff < __fmul_rn(__fmul_rn(q.grad_sq, q.sigma), lower)) q.sigma = __fmul_rn(q.sigma, scale_down);
    q.old_value = q.value;
}

__global__ void qw_set_old_value_kernel(Scal *__restrict__ sc, int ns)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s < ns) sc[s].old_value = sc[s].value;
}

// ------------------------------------------------------------------------------------------------ c-transform
// One thread per grid line: lower convex hull by the reference's stack walk (fot2d.c:66-104), dual indices by its
// merge (fot2d.c:106-127) and the dual values with its last-point guard (fot2d.c:129-147). Input element k of line
// l at in[l * in_line + k * in_elem], negated when `negate` (fot2d.c:167-169); output likewise. The hull stack of
// a line lives in global memory, entry k of line l at hull[k * nlines_total + l] (coalesced across a warp), and holds
// the point's VALUE next to its index; the two entries on top of the stack (hull walk) and the current hull edge
// (dual walk) are carried in registers. A push is then a store and nothing else, and only a pop (or a step to the
// next hull edge) reads the stack - one 8-byte load instead of the index -> value chain of dependent loads per point
// that made the 1501-point column pass 2.8 ms (now the arithmetic of the walk; same operations, same results).
__global__ void qw_dual_lines_kernel(const float *__restrict__ in, float *__restrict__ out, int2 *__restrict__ hull,
                                     int n, int nlines, int ns, int64_t rec_stride, int in_line, int in_elem, int negate)
{
    // slope grid (k + .5) / n as float and the last-point guard sp * (n - .5) / n as double depend on the position
    // only: tabulated once per block (the walk would otherwise do three double-precision divisions per point)
    extern __shared__ __align__(8) unsigned char qw_tab[];
    double *gtab = reinterpret_cast<double *>(qw_tab);              // [n] (double)sp * (n - .5) / (n * 1.0)
    float *stab = reinterpret_cast<float *>(gtab + n);              // [n] (float)((k + .5) / (n * 1.0))
    for (int k = threadIdx.x; k < n; k += blockDim.x) {
        const float sp = (float)((k + .5) / (n * 1.0));
        stab[k] = sp;
        gtab[k] = (double)sp * (n - .5) / (n * 1.0);
    }
    __syncthreads();
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int total = nlines * ns;
    if (t >= total) return;
    const int s = t / nlines, l = t - s * nlines;
    const float *u_ = in + (int64_t)s * rec_stride + (int64_t)l * in_line;
    float *d_ = out + (int64_t)s * rec_stride + (int64_t)l * in_line;
    int2 *h = hull + t;
    const float sgn = negate ? -1.f : 1.f;
#define U(k) (sgn * u_[(int64_t)(k)*in_elem])
#define HS(k, idx, val) h[(int64_t)(k)*total] = make_int2((idx), __float_as_int(val))
#define HL(k) h[(int64_t)(k)*total]
    // get_convex_hull: (ic1, u1) = top of the stack, (ic2, u2) = the entry below it
    int ic2 = 0, ic1 = 1;
    float u2 = U(0), u1 = U(1);
    HS(0, ic2, u2);
    HS(1, ic1, u1);
    int hc = 2;
    // the walk itself is a few dozen dependent instructions per point, a global load takes ~1 us: the line is read
    // QW_PF points ahead (two register chunks), so that the loads of the next chunk fly under the walk of this one
    float nxt[QW_PF];
#pragma unroll
    for (int j = 0; j < QW_PF; j++) nxt[j] = (2 + j < n) ? U(2 + j) : 0.f;
    for (int i0 = 2; i0 < n; i0 += QW_PF) {
        float cur[QW_PF];
#pragma unroll
        for (int j = 0; j < QW_PF; j++) cur[j] = nxt[j];
#pragma unroll
        for (int j = 0; j < QW_PF; j++) nxt[j] = (i0 + QW_PF + j < n) ? U(i0 + QW_PF + j) : 0.f;
#pragma unroll
        for (int j = 0; j < QW_PF; j++) {
            const int i = i0 + j;
            if (i >= n) break;
            const float ui = cur[j];
            for (;;) {
                if (hc < 2) {                             // only the first point is left: the stack becomes [first, i]
                    HS(1, i, ui);
                    hc++;
                    ic2 = ic1; u2 = u1; ic1 = i; u1 = ui;
                    break;
                }
                const float old_slope = __fdiv_rn(__fsub_rn(u1, u2), (float)(ic1 - ic2));
                const float slope = __fdiv_rn(__fsub_rn(ui, u1), (float)(i - ic1));
                if (slope >= old_slope) {
                    HS(hc, i, ui);
                    hc++;
                    ic2 = ic1; u2 = u1; ic1 = i; u1 = ui;
                    break;
                }
                hc--;                                     // pop
                ic1 = ic2; u1 = u2;
                if (hc >= 2) { const int2 e = HL(hc - 2); ic2 = e.x; u2 = __int_as_float(e.y); }
            }
        }
    }
    const float ulast = u_[(int64_t)(n - 1) * in_elem] * sgn;
    // compute_dual_indicies + compute_dual: (ic2, u2) .. (ic1, u1) = the current hull edge, entries counter-1 and counter
    int counter = 1;
    { const int2 e0 = HL(0), e1 = HL(1); ic2 = e0.x; u2 = __int_as_float(e0.y); ic1 = e1.x; u1 = __int_as_float(e1.y); }
    int2 enext = (counter + 1 < hc) ? HL(counter + 1) : make_int2(0, 0);
    float slope = __fdiv_rn(__fmul_rn((float)n, __fsub_rn(u1, u2)), (float)(ic1 - ic2));
    for (int i = 0; i < n; i++) {
        const float sp = stab[i];
        // (the reference re-evaluates the slope of the current hull edge at the top of every i; same value)
        while (sp > slope && counter < hc - 1) {
            counter++;
            ic2 = ic1; u2 = u1;
            ic1 = enext.x; u1 = __int_as_float(enext.y);
            if (counter + 1 < hc) enext = HL(counter + 1);
            slope = __fdiv_rn(__fmul_rn((float)n, __fsub_rn(u1, u2)), (float)(ic1 - ic2));
        }
        const float x = stab[ic2];                        // ic2 = hull entry counter - 1
        const float v1 = __fsub_rn(__fmul_rn(sp, x), u2);
        const float v2 = (float)(gtab[i] - (double)ulast);
        d_[(int64_t)i * in_elem] = (v1 > v2) ? v1 : v2;
    }
#undef U
#undef HS
#undef HL
}

// ------------------------------------------------------------------------------------------------ push-forward
static __device__ __forceinline__ int qw_sgn(float x) { return (x > 0) - (x < 0); }

// fot2d.c:262-288
static __device__ float qw_interp(const float *__restrict__ f, float x, float y, int n1, int n2)
{
    const int xi = (int)fmin(fmax((double)__fmul_rn(x, (float)n1) - .5, 0.), (double)(n1 - 1));
    const int yi = (int)fmin(fmax((double)__fmul_rn(y, (float)n2) - .5, 0.), (double)(n2 - 1));
    const float xfrac = (float)((double)__fsub_rn(__fmul_rn(x, (float)n1), (float)xi) - .5);
    const float yfrac = (float)((double)__fsub_rn(__fmul_rn(y, (float)n2), (float)yi) - .5);
    int xo = xi + qw_sgn(xfrac), yo = yi + qw_sgn(yfrac);
    xo = (int)fmax(fmin((double)xo, (double)(n1 - 1)), 0.);
    yo = (int)fmax(fmin((double)yo, (double)(n2 - 1)), 0.);
    const double ax = fabs((double)xfrac), ay = fabs((double)yfrac);
    const float v1 = (float)((1 - ax) * (1 - ay) * (double)f[yi * n1 + xi]);
    const float v2 = (float)(ax * (1 - ay) * (double)f[yi * n1 + xo]);
    const float v3 = (float)((1 - ax) * ay * (double)f[yo * n1 + xi]);
    const float v4 = (float)(ax * ay * (double)f[yo * n1 + xo]);
    return __fadd_rn(__fadd_rn(__fadd_rn(v1, v2), v3), v4);
}

// fot2d.c:290-322: centred differences of the bilinear interpolant at the (n1+1) x (n2+1) cell corners
__global__ void qw_map_kernel(Ptrs p, const float *__restrict__ pot, int n1, int n2, int ns)
{
    const int mc = (n1 + 1) * (n2 + 1);
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (int64_t)mc * ns) return;
    const int s = (int)(t / mc), k = (int)(t - (int64_t)s * mc);
    const int i = k / (n1 + 1), j = k - i * (n1 + 1);
    const float *f = pot + (int64_t)s * n1 * n2;
    const float xstep = (float)(1.0 / n1), ystep = (float)(1.0 / n2);
    const float x = (float)(j / (n1 * 1.0)), y = (float)(i / (n2 * 1.0));
    const float dxp = qw_interp(f, __fadd_rn(x, xstep), y, n1, n2), dxm = qw_interp(f, __fsub_rn(x, xstep), y, n1, n2);
    const float dyp = qw_interp(f, x, __fadd_rn(y, ystep), n1, n2), dym = qw_interp(f, x, __fsub_rn(y, ystep), n1, n2);
    p.xmap[t] = (float)(.5 * n1 * (double)__fsub_rn(dxp, dxm));
    p.ymap[t] = (float)(.5 * n2 * (double)__fsub_rn(dyp, dym));
}

// fot2d.c:398-459: every cell of positive mass is sampled xSamples x ySamples times and splatted bilinearly at the
// mapped positions. Accumulation in 2^-40 fixed point (order independent); one thread per cell.
__global__ void qw_splat_kernel(Ptrs p, const float *__restrict__ dens, int n1, int n2, int ns)
{
    const int pc = n1 * n2, mc = (n1 + 1) * (n2 + 1);
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (int64_t)pc * ns) return;
    const int s = (int)(t / pc), k = (int)(t - (int64_t)s * pc);
    if (p.sc[s].skip) return;
    const float mass = dens[t];
    if (!(mass > 0)) return;
    const int i = k / n1, j = k - i * n1;
    const float *xm = p.xmap + (int64_t)s * mc, *ym = p.ymap + (int64_t)s * mc;
    long long *acc = p.racc + (int64_t)s * pc;
    const float x00 = xm[i * (n1 + 1) + j], x01 = xm[i * (n1 + 1) + j + 1];
    const float x10 = xm[(i + 1) * (n1 + 1) + j], x11 = xm[(i + 1) * (n1 + 1) + j + 1];
    const float y00 = ym[i * (n1 + 1) + j], y01 = ym[i * (n1 + 1) + j + 1];
    const float y10 = ym[(i + 1) * (n1 + 1) + j], y11 = ym[(i + 1) * (n1 + 1) + j + 1];
    const float xcut = (float)pow(1.0 / n1, 1.0 / 3), ycut = (float)pow(1.0 / n2, 1.0 / 3);
    const float xs = fmaxf(fabsf(__fsub_rn(x01, x00)), fabsf(__fsub_rn(x11, x10)));
    const float ys = fmaxf(fabsf(__fsub_rn(y10, y00)), fabsf(__fsub_rn(y11, y01)));
    if (!(xs < xcut && ys < ycut)) return;
    const int nxs = (int)(2 * fmax((double)__fmul_rn((float)n1, xs), 1.));
    const int nys = (int)(2 * fmax((double)__fmul_rn((float)n2, ys), 1.));
    const float factor = (float)(1 / (nxs * nys * 1.0));
    for (int l = 0; l < nys; l++) {
        const float b = (float)((l + .5) / (nys * 1.0));
        for (int q = 0; q < nxs; q++) {
            const float a = (float)((q + .5) / (nxs * 1.0));
            const float w00 = __fmul_rn(__fsub_rn(1.f, b), __fsub_rn(1.f, a)), w01 = __fmul_rn(__fsub_rn(1.f, b), a);
            const float w10 = __fmul_rn(b, __fsub_rn(1.f, a)), w11 = __fmul_rn(a, b);
            const float xp = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(w00, x00), __fmul_rn(w01, x01)), __fmul_rn(w10, x10)),
                                       __fmul_rn(w11, x11));
            const float yp = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(w00, y00), __fmul_rn(w01, y01)), __fmul_rn(w10, y10)),
                                       __fmul_rn(w11, y11));
            const float X = (float)((double)__fmul_rn(xp, (float)n1) - .5), Y = (float)((double)__fmul_rn(yp, (float)n2) - .5);
            int xi = (int)X, yi = (int)Y;
            const float xf = __fsub_rn(X, (float)xi), yf = __fsub_rn(Y, (float)yi);
            int xo = xi + 1, yo = yi + 1;
            xi = min(max(xi, 0), n1 - 1); xo = min(max(xo, 0), n1 - 1);
            yi = min(max(yi, 0), n2 - 1); yo = min(max(yo, 0), n2 - 1);
            const float c00 = __fmul_rn(__fmul_rn(__fmul_rn(__fsub_rn(1.f, xf), __fsub_rn(1.f, yf)), mass), factor);
            const float c10 = __fmul_rn(__fmul_rn(__fmul_rn(__fsub_rn(1.f, xf), yf), mass), factor);
            const float c01 = __fmul_rn(__fmul_rn(__fmul_rn(xf, __fsub_rn(1.f, yf)), mass), factor);
            const float c11 = __fmul_rn(__fmul_rn(__fmul_rn(xf, yf), mass), factor);
            atomicAdd((unsigned long long *)&acc[yi * n1 + xi], (unsigned long long)__float2ll_rn(c00 * QW_FIXED_SCALE));
            atomicAdd((unsigned long long *)&acc[yo * n1 + xi], (unsigned long long)__float2ll_rn(c10 * QW_FIXED_SCALE));
            atomicAdd((unsigned long long *)&acc[yi * n1 + xo], (unsigned long long)__float2ll_rn(c01 * QW_FIXED_SCALE));
            atomicAdd((unsigned long long *)&acc[yo * n1 + xo], (unsigned long long)__float2ll_rn(c11 * QW_FIXED_SCALE));
        }
    }
}

__global__ void qw_fixed_to_float_kernel(Ptrs p, int64_t n)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p.rho[i] = (float)((double)p.racc[i] * QW_FIXED_INV);
}

// rho *= totalMass / sum   (fot2d.c:472-474, totalMass = 1)
__global__ void qw_rho_scale_kernel(Ptrs p, int pc, int ns)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)pc * ns) return;
    const Scal &q = p.sc[i / pc];
    if (q.skip) return;
    p.rho[i] = __fmul_rn(p.rho[i], __fdiv_rn(1.f, q.rho_sum));
}

// fot2d.c:597-604: potentials -> Kantorovich potentials
__global__ void qw_finish_potentials_kernel(Ptrs p, int n1, int n2, int ns)
{
    const int pc = n1 * n2;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)pc * ns) return;
    const int k = (int)(i % pc);
    const int r = k / n1, j = k - r * n1;
    const float x = (float)((j + .5) / (n1 * 1.0)), y = (float)((r + .5) / (n2 * 1.0));
    const double z = .5 * (double)__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y));
    p.phi[i] = (float)(z - (double)p.phi[i]);
    p.dual[i] = (float)(z - (double)p.dual[i]);
}

// fot2d.c:646-653 (grad initialised to 1, w2.c:32-33) and misfit.py:79,104: adjoint source = grad / mass
__global__ void qw_adjoint_kernel(Ptrs p, float *__restrict__ adj, double *__restrict__ fval, int pc, int ns)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)pc * ns) return;
    const Scal &q = p.sc[i / pc];
    float g = 1.f;
    if (!q.skip) g = __fmul_rn(1.f, __fdiv_rn(__fsub_rn(p.dual[i], q.term), q.sum1));
    adj[i] = __fdiv_rn(g, q.mass);
    (void)fval;
}

__global__ void qw_loss_kernel(Ptrs p, int ns, double *__restrict__ fval, float *__restrict__ loss_out)
{
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    double f = 0.0;
    for (int s = 0; s < ns; s++) {
        const float l = p.sc[s].skip ? 0.f : p.sc[s].old_value;      // compute_l2_fot2d returns oldValue (fot2d.c:606)
        p.sc[s].loss = l;
        if (loss_out) loss_out[s] = l;
        f += (double)l;
    }
    fval[0] += f;
}

// ------------------------------------------------------------------------------------------------ DCTs (cuFFT)
// FFTW_REDFT10 / FFTW_REDFT01 along one dimension through a real FFT of the same length (Makhoul's reordering):
//   DCT-II : v[n] = x[2n], v[N-1-n] = x[2n+1];  V = FFT(v);  Y[k] = 2 Re(e^{-i pi k / 2N} V[k])
//   DCT-III: V[k] = e^{i pi k / 2N} (X[k] - i X[N-k]), X[N] = 0;  v = N * IFFT(V);  y[2n] = v[n], y[2n+1] = v[N-1-n]
//            (the complex-to-real inverse of cuFFT is not used: see qw_dct3_pre)
// Lines are gathered into / scattered from contiguous cuFFT buffers, which also takes care of the slow dimension.
__global__ void qw_dct2_pre(const float *__restrict__ in, float *__restrict__ v, int N, int nlines, int ns, int64_t rec_stride,
                            int line_stride, int elem_stride)
{
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (int64_t)N * nlines * ns) return;
    const int n = (int)(t % N);
    const int64_t ln = t / N;
    const int s = (int)(ln / nlines), l = (int)(ln - (int64_t)s * nlines);
    const int src = (n < (N + 1) / 2) ? 2 * n : 2 * (N - 1 - n) + 1;
    v[t] = in[(int64_t)s * rec_stride + (int64_t)l * line_stride + (int64_t)src * elem_stride];
}

__global__ void qw_dct2_post(const float2 *__restrict__ V, float *__restrict__ out, int N, int nlines, int ns, int64_t rec_stride,
                             int line_stride, int elem_stride)
{
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (int64_t)N * nlines * ns) return;
    const int k = (int)(t % N);
    const int64_t ln = t / N;
    const int s = (int)(ln / nlines), l = (int)(ln - (int64_t)s * nlines);
    const int nc = N / 2 + 1;
    float2 c = (k < nc) ? V[ln * nc + k] : V[ln * nc + (N - k)];
    if (k >= nc) c.y = -c.y;
    double sn, cs;
    sincospi((double)k / (2.0 * N), &sn, &cs);
    out[(int64_t)s * rec_stride + (int64_t)l * line_stride + (int64_t)k * elem_stride] =
        (float)(2.0 * (cs * (double)c.x + sn * (double)c.y));
}

// The inverse also goes through a REAL-to-complex FFT: the Hermitian spectrum V (V[N-k] = conj V[k]) is folded into the
// real sequence w[k] = Re V[k] - Im V[k]; then sum_k V[k] e^{+2 pi i nk/N} = Re W[n] - Im W[n] with W = FFT(w).
__global__ void qw_dct3_pre(const float *__restrict__ in, float *__restrict__ w, int N, int nlines, int ns, int64_t rec_stride,
                            int line_stride, int elem_stride)
{
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (int64_t)N * nlines * ns) return;
    const int k = (int)(t % N);
    const int64_t ln = t / N;
    const int s = (int)(ln / nlines), l = (int)(ln - (int64_t)s * nlines);
    const float *x = in + (int64_t)s * rec_stride + (int64_t)l * line_stride;
    const int kk = (k <= N / 2) ? k : N - k;
    const double a = (double)x[(int64_t)kk * elem_stride];
    const double b = (kk > 0) ? (double)x[(int64_t)(N - kk) * elem_stride] : 0.0;
    double sn, cs;
    sincospi((double)kk / (2.0 * N), &sn, &cs);
    const double A = cs * a + sn * b;
    double B = sn * a - cs * b;
    if (k > N / 2) B = -B;
    w[t] = (float)(A - B);
}

__global__ void qw_dct3_post(const float2 *__restrict__ W, float *__restrict__ out, int N, int nlines, int ns, int64_t rec_stride,
                             int line_stride, int elem_stride)
{
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (int64_t)N * nlines * ns) return;
    const int j = (int)(t % N);
    const int64_t ln = t / N;
    const int s = (int)(ln / nlines), l = (int)(ln - (int64_t)s * nlines);
    // y[2h] = v[h], y[2h+1] = v[N-1-h]. Written without a select on purpose: for `(j & 1) ? N - 1 - (j - 1) / 2 : j / 2`
    // nvcc 12.9 emits LEA.HI.SX32 on -j, which returned N - (j - 1) / 2 on sm_100a (odd outputs shifted by one).
    const int h = (int)((unsigned)j >> 1);
    const int n = h + (j & 1) * (N - 1 - 2 * h);
    const int nc = N / 2 + 1;
    float v;
    if (n < nc) { const float2 c = W[ln * nc + n]; v = (float)((double)c.x - (double)c.y); }
    else { const float2 c = W[ln * nc + (N - n)]; v = (float)((double)c.x + (double)c.y); }
    out[(int64_t)s * rec_stride + (int64_t)l * line_stride + (int64_t)j * elem_stride] = v;
}

// cuFFT plans are created once per (length, batch) and live for the process (cuFFT allocates their work areas).
struct PlanKey {
    int n, batch, kind;
    bool operator<(const PlanKey &o) const { return std::tie(n, batch, kind) < std::tie(o.n, o.batch, o.kind); }
};
static std::map<PlanKey, cufftHandle> g_plans;
static std::mutex g_plan_mutex;

static int get_plan(int n, int batch, cufftType type, cufftHandle *out)
{
    std::lock_guard<std::mutex> lock(g_plan_mutex);
    const PlanKey key{n, batch, (int)type};
    auto it = g_plans.find(key);
    if (it != g_plans.end()) { *out = it->second; return 0; }
    cufftHandle h;
    int dims[1] = {n};
    const cufftResult r = cufftPlanMany(&h, 1, dims, nullptr, 1, 0, nullptr, 1, 0, type, batch);
    if (r != CUFFT_SUCCESS) {
        set_error("cufftPlanMany(n=%d, batch=%d) failed with %d", n, batch, (int)r);
        return B2FWI_ECUDA;
    }
    g_plans[key] = h;
    *out = h;
    return 0;
}

#define QW_GRID(n) (unsigned)(((n) + 255) / 256), 256, 0, st

// 2-D REDFT10 (forward = true) or REDFT01 of ws, in place, all records
static int g_dct_only = -1;      // diagnostics: restrict dct2d to one dimension (b2fwi_qw2d_debug_step ops 6, 7)

static int dct2d(const Ptrs &p, int n1, int n2, int ns, bool forward, cudaStream_t st)
{
    const int pc = n1 * n2;
    const int64_t tot = (int64_t)pc * ns;
    // dimension n1: lines = rows (stride n1, elements contiguous); dimension n2: lines = columns (stride 1, elements n1 apart)
    const int N[2] = {n1, n2}, NL[2] = {n2, n1}, LS[2] = {n1, 1}, ES[2] = {1, n1};
    for (int pass = 0; pass < 2; pass++) {
        const int d = forward ? pass : 1 - pass;       // the inverse runs the dimensions in the opposite order
        if (g_dct_only >= 0 && d != g_dct_only) continue;
        cufftHandle h;
        int rc = get_plan(N[d], NL[d] * ns, CUFFT_R2C, &h);
        if (rc) return rc;
        if (cufftSetStream(h, st) != CUFFT_SUCCESS) { set_error("cufftSetStream failed"); return B2FWI_ECUDA; }
        if (forward) {
            qw_dct2_pre<<<QW_GRID(tot)>>>(p.ws, p.freal, N[d], NL[d], ns, pc, LS[d], ES[d]);
            if (cufftExecR2C(h, p.freal, (cufftComplex *)p.fcplx) != CUFFT_SUCCESS) { set_error("cufftExecR2C failed"); return B2FWI_ECUDA; }
            qw_dct2_post<<<QW_GRID(tot)>>>(p.fcplx, p.ws, N[d], NL[d], ns, pc, LS[d], ES[d]);
        } else {
            qw_dct3_pre<<<QW_GRID(tot)>>>(p.ws, p.freal, N[d], NL[d], ns, pc, LS[d], ES[d]);
            if (cufftExecR2C(h, p.freal, (cufftComplex *)p.fcplx) != CUFFT_SUCCESS) { set_error("cufftExecR2C failed"); return B2FWI_ECUDA; }
            qw_dct3_post<<<QW_GRID(tot)>>>(p.fcplx, p.ws, N[d], NL[d], ns, pc, LS[d], ES[d]);
        }
        count_launch(3);
    }
    B2_CUDA(cudaGetLastError());
    return 0;
}

// fot2d.c:479-503
static int update_potential(const Ptrs &p, float *pot, const float *other, int n1, int n2, int ns, cudaStream_t st)
{
    const int pc = n1 * n2;
    const int64_t tot = (int64_t)pc * ns;
    qw_sub_kernel<<<QW_GRID(tot)>>>(p.rho, other, p.ws, tot);
    int rc = dct2d(p, n1, n2, ns, true, st);
    if (rc) return rc;
    qw_poisson_scale_kernel<<<QW_GRID(tot)>>>(p.ws, p.kern, pc, ns);
    if ((rc = dct2d(p, n1, n2, ns, false, st))) return rc;
    qw_axpy_kernel<<<QW_GRID(tot)>>>(pot, p.ws, p.sc, pc, ns);
    qw_seq_kernel<SEQ_H1><<<ns, 512, 0, st>>>(p, n1, n2, other);
    count_launch(4);
    return 0;
}

// fot2d.c:157-183: dual = c-transform(u): rows, transpose + negate, columns, transpose back (the transposes are strides here)
static void dual2d(const Ptrs &p, const float *u, float *dual, int n1, int n2, int ns, cudaStream_t st)
{
    const int pc = n1 * n2;
    // a line is one long dependent chain: with few lines (the column pass: nrec * nshots) one warp per block spreads
    // them over all SMs instead of four warps on a third of them
    const int bs1 = (n2 * ns >= 148 * 128) ? 128 : 32, bs2 = (n1 * ns >= 148 * 128) ? 128 : 32;
    static bool configured = false;          // position tables: 12 B per point of a line (records of up to ~18 k samples)
    if (!configured) {
        cudaFuncSetAttribute(qw_dual_lines_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
        configured = true;
    }
    qw_dual_lines_kernel<<<(unsigned)((n2 * ns + bs1 - 1) / bs1), bs1, (size_t)n1 * 12, st>>>(u, p.tmp, p.hull, n1, n2, ns, pc, n1, 1, 0);
    qw_dual_lines_kernel<<<(unsigned)((n1 * ns + bs2 - 1) / bs2), bs2, (size_t)n2 * 12, st>>>(p.tmp, dual, p.hull, n2, n1, ns, pc, 1, n1, 1);
    count_launch(2);
}

static void pushforward(const Ptrs &p, const float *pot, const float *dens, int n1, int n2, int ns, cudaStream_t st)
{
    const int pc = n1 * n2, mc = (n1 + 1) * (n2 + 1);
    const int64_t tot = (int64_t)pc * ns;
    qw_map_kernel<<<QW_GRID((int64_t)mc * ns)>>>(p, pot, n1, n2, ns);
    cudaMemsetAsync(p.racc, 0, sizeof(long long) * tot, st);
    qw_splat_kernel<<<QW_GRID(tot)>>>(p, dens, n1, n2, ns);
    qw_fixed_to_float_kernel<<<QW_GRID(tot)>>>(p, tot);
    qw_seq_kernel<SEQ_RHO><<<ns, 512, 0, st>>>(p, n1, n2, nullptr);
    qw_rho_scale_kernel<<<QW_GRID(tot)>>>(p, pc, ns);
    count_launch(5);
}

static size_t align_up(size_t x) { return (x + 255) & ~(size_t)255; }

struct Layout2 {
    size_t field, maps, hull, racc, freal, fcplx, scal, kern, total;
};

static Layout2 layout(int n1, int n2, int ns)
{
    Layout2 L;
    const size_t pc = (size_t)n1 * n2, mc = (size_t)(n1 + 1) * (n2 + 1);
    L.field = align_up(pc * ns * sizeof(float));
    L.maps = align_up(mc * ns * sizeof(float));
    L.hull = align_up(pc * ns * sizeof(int2));
    L.racc = align_up(pc * ns * sizeof(long long));
    L.freal = align_up(pc * ns * sizeof(float));
    const size_t c1 = (size_t)(n1 / 2 + 1) * n2, c2 = (size_t)(n2 / 2 + 1) * n1;
    L.fcplx = align_up((c1 > c2 ? c1 : c2) * ns * sizeof(float2));
    L.scal = align_up((size_t)ns * sizeof(Scal));
    L.kern = align_up(pc * sizeof(float));
    L.total = 7 * L.field + 2 * L.maps + L.hull + L.racc + L.freal + L.fcplx + L.scal + L.kern;
    return L;
}

}  // namespace qw
}  // namespace b2fwi

using namespace b2fwi;
using namespace b2fwi::qw;

extern "C" {

int64_t b2fwi_qw2d_scratch_bytes(int32_t nt, int32_t nrec, int32_t nshots)
{
    if (nt < 2 || nrec < 2 || nshots < 1) return 0;
    return (int64_t)layout(nrec, nt, nshots).total;
}

// Single solver steps on caller-provided single-record fields (diagnostics / stage-by-stage parity tests):
//   op 0  dual = c-transform(a)                                   (compute_2d_dual)
//   op 1  a += sigma * Poisson(b - c); scal[0] = H^-1 residual    (update_potential; a = potential, b = rho, c = other)
//   op 2  dual = push-forward of density b by the gradient of a   (calc_pushforward_map + sampling_pushforward)
//   op 3  scal[0] = W2 value of (phi = a, dual = b, mu = c, nu = d)
int b2fwi_qw2d_debug_step(int32_t op, int32_t nt, int32_t nrec, float *a, float *b_, float *c, float *d, float sigma,
                          float *out, float *scal, void *scratch, void *stream)
{
    B2_CHECK_ARG(scratch && a && nt >= 2 && nrec >= 2, "bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    const int n1 = nrec, n2 = nt, ns = 1, pc = n1 * n2;
    const Layout2 L = layout(n1, n2, ns);
    char *b = reinterpret_cast<char *>(scratch);
    Ptrs p;
    float **fields[7] = {&p.mu, &p.nu, &p.phi, &p.dual, &p.rho, &p.ws, &p.tmp};
    for (int k = 0; k < 7; k++) { *fields[k] = reinterpret_cast<float *>(b); b += L.field; }
    p.xmap = reinterpret_cast<float *>(b); b += L.maps;
    p.ymap = reinterpret_cast<float *>(b); b += L.maps;
    p.hull = reinterpret_cast<int2 *>(b); b += L.hull;
    p.racc = reinterpret_cast<long long *>(b); b += L.racc;
    p.freal = reinterpret_cast<float *>(b); b += L.freal;
    p.fcplx = reinterpret_cast<float2 *>(b); b += L.fcplx;
    p.sc = reinterpret_cast<Scal *>(b); b += L.scal;
    p.kern = reinterpret_cast<float *>(b);
    Scal h;
    memset(&h, 0, sizeof(h));
    h.sigma = sigma;
    h.sum1 = 1.f;
    B2_CUDA(cudaMemcpyAsync(p.sc, &h, sizeof(h), cudaMemcpyHostToDevice, st));
    int rc = 0;
    if (op == 0) {
        dual2d(p, a, out, n1, n2, ns, st);
    } else if (op == 1) {
        qw_kernel_kernel<<<QW_GRID(pc)>>>(p.kern, n1, n2);
        B2_CUDA(cudaMemcpyAsync(p.rho, b_, sizeof(float) * pc, cudaMemcpyDeviceToDevice, st));
        rc = update_potential(p, a, c, n1, n2, ns, st);
        B2_CUDA(cudaMemcpyAsync(scal, &p.sc->grad_sq, sizeof(float), cudaMemcpyDeviceToDevice, st));
    } else if (op == 2) {
        pushforward(p, a, b_, n1, n2, ns, st);
        B2_CUDA(cudaMemcpyAsync(out, p.rho, sizeof(float) * pc, cudaMemcpyDeviceToDevice, st));
    } else if (op == 3) {
        B2_CUDA(cudaMemcpyAsync(p.phi, a, sizeof(float) * pc, cudaMemcpyDeviceToDevice, st));
        B2_CUDA(cudaMemcpyAsync(p.dual, b_, sizeof(float) * pc, cudaMemcpyDeviceToDevice, st));
        B2_CUDA(cudaMemcpyAsync(p.mu, c, sizeof(float) * pc, cudaMemcpyDeviceToDevice, st));
        B2_CUDA(cudaMemcpyAsync(p.nu, d, sizeof(float) * pc, cudaMemcpyDeviceToDevice, st));
        qw_seq_kernel<SEQ_W2><<<1, 256, 0, st>>>(p, n1, n2, nullptr);
        B2_CUDA(cudaMemcpyAsync(scal, &p.sc->value, sizeof(float), cudaMemcpyDeviceToDevice, st));
    } else if (op == 8 || op == 9) {          // diagnostics: w = dct3_pre(a) along n1 (8); W = FFT(w) as interleaved floats (9)
        B2_CUDA(cudaMemcpyAsync(p.ws, a, sizeof(float) * pc, cudaMemcpyDeviceToDevice, st));
        qw_dct3_pre<<<QW_GRID(pc)>>>(p.ws, p.freal, n1, n2, 1, pc, n1, 1);
        cufftHandle h;
        rc = get_plan(n1, n2, CUFFT_R2C, &h);
        if (rc) return rc;
        cufftSetStream(h, st);
        if (op == 9) {
            if (cufftExecR2C(h, p.freal, (cufftComplex *)p.fcplx) != CUFFT_SUCCESS) { set_error("cufftExecR2C failed"); return B2FWI_ECUDA; }
            B2_CUDA(cudaMemcpyAsync(out, p.fcplx, sizeof(float) * pc, cudaMemcpyDeviceToDevice, st));
        } else {
            B2_CUDA(cudaMemcpyAsync(out, p.freal, sizeof(float) * pc, cudaMemcpyDeviceToDevice, st));
        }
    } else if (op == 6 || op == 7) {          // out = REDFT01 of a along n1 only (6) / along n2 only (7)
        B2_CUDA(cudaMemcpyAsync(p.ws, a, sizeof(float) * pc, cudaMemcpyDeviceToDevice, st));
        g_dct_only = op - 6;
        rc = dct2d(p, n1, n2, ns, false, st);
        g_dct_only = -1;
        B2_CUDA(cudaMemcpyAsync(out, p.ws, sizeof(float) * pc, cudaMemcpyDeviceToDevice, st));
    } else if (op == 4 || op == 5) {          // out = REDFT10 (4) / REDFT01 (5) of a, both dimensions
        B2_CUDA(cudaMemcpyAsync(p.ws, a, sizeof(float) * pc, cudaMemcpyDeviceToDevice, st));
        rc = dct2d(p, n1, n2, ns, op == 4, st);
        B2_CUDA(cudaMemcpyAsync(out, p.ws, sizeof(float) * pc, cudaMemcpyDeviceToDevice, st));
    } else {
        set_error("unknown debug op %d", op);
        return B2FWI_EINVAL;
    }
    B2_CUDA(cudaGetLastError());
    return rc;
}

int b2fwi_qw2d_misfit(const float *syn, const float *obs, const float *dw, int32_t nt, int32_t nrec, int32_t nshots,
                      double gamma, int32_t num_steps, float step_scale, float *adjsrc_out, double *fval_out,
                      float *loss_out, void *scratch, void *stream)
{
    B2_CHECK_ARG(syn && obs && adjsrc_out && fval_out && scratch, "NULL argument");
    B2_CHECK_ARG(nt >= 2 && nrec >= 2 && nshots >= 1 && num_steps >= 0, "bad sizes nt=%d nrec=%d nshots=%d", nt, nrec, nshots);
    B2_CHECK_ARG((int64_t)nt * nrec * nshots < (1ll << 31) / 4, "records too large for 32-bit indexing");
    cudaStream_t st = (cudaStream_t)stream;
    const int n1 = nrec, n2 = nt, ns = nshots, pc = n1 * n2;
    const int64_t tot = (int64_t)pc * ns;
    const Layout2 L = layout(n1, n2, ns);
    char *b = reinterpret_cast<char *>(scratch);
    Ptrs p;
    float **fields[7] = {&p.mu, &p.nu, &p.phi, &p.dual, &p.rho, &p.ws, &p.tmp};
    for (int k = 0; k < 7; k++) { *fields[k] = reinterpret_cast<float *>(b); b += L.field; }
    p.xmap = reinterpret_cast<float *>(b); b += L.maps;
    p.ymap = reinterpret_cast<float *>(b); b += L.maps;
    p.hull = reinterpret_cast<int2 *>(b); b += L.hull;
    p.racc = reinterpret_cast<long long *>(b); b += L.racc;
    p.freal = reinterpret_cast<float *>(b); b += L.freal;
    p.fcplx = reinterpret_cast<float2 *>(b); b += L.fcplx;
    p.sc = reinterpret_cast<Scal *>(b); b += L.scal;
    p.kern = reinterpret_cast<float *>(b);

    // misfit.py:18-45,73 and normalize.c
    qw_min_kernel<<<ns, 256, 0, st>>>(syn, obs, dw, pc, gamma, p.sc);
    qw_shift_kernel<<<QW_GRID(tot)>>>(syn, obs, dw, pc, ns, p.sc, p.mu, p.nu);
    qw_mass_kernel<<<ns, 1 << MASS_DEPTH, 0, st>>>(p.mu, pc, ns, p.sc);
    qw_seq_kernel<SEQ_SUM_F><<<ns, 512, 0, st>>>(p, n1, n2, nullptr);
    qw_seq_kernel<SEQ_SUM_G><<<ns, 512, 0, st>>>(p, n1, n2, nullptr);
    qw_normalize_kernel<<<QW_GRID(tot)>>>(p, n1, n2, ns);
    qw_sigma_kernel<<<ns, 256, 0, st>>>(p, pc, step_scale);
    qw_kernel_kernel<<<QW_GRID(pc)>>>(p.kern, n1, n2);
    qw_seq_kernel<SEQ_W2><<<ns, 512, 0, st>>>(p, n1, n2, nullptr);          // oldValue (fot2d.c:534)
    qw_set_old_value_kernel<<<(ns + 31) / 32, 32, 0, st>>>(p.sc, ns);
    count_launch(10);
    B2_CUDA(cudaGetLastError());

    int rc;
    for (int it = 0; it < num_steps; it++) {                                // fot2d.c:541-596
        if ((rc = update_potential(p, p.phi, p.nu, n1, n2, ns, st))) return rc;
        dual2d(p, p.phi, p.dual, n1, n2, ns, st);                           // convexify(phi, dual)
        dual2d(p, p.dual, p.phi, n1, n2, ns, st);
        qw_seq_kernel<SEQ_W2><<<ns, 512, 0, st>>>(p, n1, n2, nullptr);
        qw_step_update_kernel<<<(ns + 31) / 32, 32, 0, st>>>(p.sc, ns);
        pushforward(p, p.phi, p.nu, n1, n2, ns, st);
        if ((rc = update_potential(p, p.dual, p.mu, n1, n2, ns, st))) return rc;
        dual2d(p, p.dual, p.phi, n1, n2, ns, st);                           // convexify(dual, phi)
        dual2d(p, p.phi, p.dual, n1, n2, ns, st);
        pushforward(p, p.dual, p.mu, n1, n2, ns, st);
        qw_seq_kernel<SEQ_W2><<<ns, 512, 0, st>>>(p, n1, n2, nullptr);
        qw_step_update_kernel<<<(ns + 31) / 32, 32, 0, st>>>(p.sc, ns);
        count_launch(4);
        B2_CUDA(cudaGetLastError());
    }
    qw_finish_potentials_kernel<<<QW_GRID(tot)>>>(p, n1, n2, ns);
    qw_seq_kernel<SEQ_TERM><<<ns, 512, 0, st>>>(p, n1, n2, nullptr);
    qw_adjoint_kernel<<<QW_GRID(tot)>>>(p, adjsrc_out, fval_out, pc, ns);
    qw_loss_kernel<<<1, 32, 0, st>>>(p, ns, fval_out, loss_out);
    count_launch(4);
    B2_CUDA(cudaGetLastError());
    return 0;
}

}  // extern "C"

// K1: the s-step matrix powers kernel (fp64).  One halo exchange, then s SpMV steps back to back with the
// Newton shift (and the conjugate-pair term) fused into the SpMV epilogue.
//
// Reference counterparts: SpMV.m:6-8, matrix_powers_monomial.m:6-12, matrix_powers_newton.m:15-54.
// Arithmetic order follows the reference: the full row sum first (ascending column order), then
// ``w - re(l_k)*x_i`` as a rounded product and a rounded subtraction, then ``+ im(l_k)^2*xprev_i``.
#include <algorithm>

#include "matrix.h"

using namespace calz;

namespace {

constexpr int kSpmvThreads = 256;

__device__ __forceinline__ double newton_epilogue(double w, double xi, double xp, double shift, double pair) {
    // matrix_powers_newton.m:34,40-41,43 -- two (three) separately rounded operations, no contraction
    double r = __dsub_rn(w, __dmul_rn(shift, xi));
    if (pair != 0.0) r = __dadd_rn(r, __dmul_rn(pair, xp));
    return r;
}

// ---- SELL-32-sigma: one warp per slice, one lane per row, column-major slice => fully coalesced A stream
template <bool NEWTON>
__global__ void __launch_bounds__(kSpmvThreads)
k_spmv_sell(const int32_t* __restrict__ slice_ptr, const int32_t* __restrict__ col, const double* __restrict__ val,
            const int32_t* __restrict__ perm, const double* __restrict__ x, const double* __restrict__ xprev,
            double* __restrict__ y, int64_t slice_lo, int64_t slice_hi, int64_t n_loc, double shift, double pair) {
    const int lane = threadIdx.x & 31;
    const int64_t slice = slice_lo + (int64_t)blockIdx.x * (kSpmvThreads / 32) + (threadIdx.x >> 5);
    if (slice >= slice_hi) return;
    const int32_t p0 = __ldg(slice_ptr + slice), p1 = __ldg(slice_ptr + slice + 1);
    const int64_t base = (int64_t)p0 * 32 + lane;
    const double* __restrict__ v = val + base;
    const int32_t* __restrict__ c = col + base;
    const int w = p1 - p0;
    double sum = 0.0;
    int j = 0;
    for (; j + 4 <= w; j += 4) {
        const double a0 = __ldg(v + (j + 0) * 32), a1 = __ldg(v + (j + 1) * 32);
        const double a2 = __ldg(v + (j + 2) * 32), a3 = __ldg(v + (j + 3) * 32);
        const int32_t c0 = __ldg(c + (j + 0) * 32), c1 = __ldg(c + (j + 1) * 32);
        const int32_t c2 = __ldg(c + (j + 2) * 32), c3 = __ldg(c + (j + 3) * 32);
        const double x0 = x[c0], x1 = x[c1], x2 = x[c2], x3 = x[c3];
        sum = fma(a0, x0, sum);
        sum = fma(a1, x1, sum);
        sum = fma(a2, x2, sum);
        sum = fma(a3, x3, sum);
    }
    for (; j < w; ++j) sum = fma(__ldg(v + j * 32), x[__ldg(c + j * 32)], sum);
    const int64_t r = slice * 32 + lane;
    if (r < n_loc) {
        const int64_t row = perm ? (int64_t)__ldg(perm + r) : r;
        if (NEWTON) sum = newton_epilogue(sum, x[row], pair != 0.0 ? xprev[row] : 0.0, shift, pair);
        y[row] = sum;
    }
}

// ---- dictionary-coded SELL: one warp per slice, one lane per row; each lane pulls 8 code bytes per 64-bit load and looks
//      {value, column offset} up in a <= 4 KB dictionary.  Same summation order as CSR/SELL (bit-identical).
//      Where the dictionary lives is a template parameter (all three measured on B200):
//        DM_SHARED  : split value / byte-offset arrays in shared memory (LDS.64 + LDS.32 per non-zero): 0.137 ms per C3 SpMV
//        DM_CONST   : kernel parameter = constant bank (LDC): a warp-uniform code costs no L1 data-pipe wavefront, 0.121 ms
//      (one 16-byte shared entry read with LDS.128 was measured too: 0.131 ms, dropped)
constexpr int kDictSlicesPerWarp = 2;     // independent slices per warp: the loads of both are in flight together
enum { DM_SHARED = 0, DM_CONST = 2 };

struct __align__(16) DictEnt {
    double v;
    int offb;                             // column offset pre-scaled to bytes
    int pad;
};
struct DictParam {                        // 4 KB, passed by value as a __grid_constant__ kernel parameter
    DictEnt e[256];
};

template <int DM>
__device__ __forceinline__ uint32_t dict_stage(unsigned char* sraw, const DictParam& P) {
    if (DM == DM_CONST) return 0;
    uint32_t sbase;
    {   // opaque copy: keeps the base in a register instead of re-deriving it (S2R + LEA) at every use
        const uint32_t t = (uint32_t)__cvta_generic_to_shared(sraw);
        asm volatile("mov.u32 %0, %1;" : "=r"(sbase) : "r"(t));
    }
    for (int i = threadIdx.x; i < 256; i += kSpmvThreads) {
        *reinterpret_cast<double*>(sraw + 8 * i) = P.e[i].v;
        *reinterpret_cast<int*>(sraw + 2048 + 4 * i) = P.e[i].offb;
    }
    __syncthreads();
    return sbase;
}

// the 8 non-zeros of one code word pair, ascending column order.  `sbase` is the 4 KB-aligned shared-window address of the
// dictionary, so an entry address is one LOP3: (bits & mask) | sbase.
template <int DM, int NS>
__device__ __forceinline__ void dict_fma8(const DictParam& P, uint32_t sbase, const uint2 (&w)[NS], const char* const (&xr)[NS],
                                          double (&sum)[NS]) {
#pragma unroll
    for (int q = 0; q < 8; ++q) {
#pragma unroll
        for (int i = 0; i < NS; ++i) {
            const unsigned int half = q < 4 ? w[i].x : w[i].y;
            // the code byte already scaled to the 16-byte entry stride: (c << 4) in one shift + one mask
            const unsigned int c16 = ((q & 3) == 0 ? (half << 4) : (half >> (8 * (q & 3) - 4))) & 0xff0u;
            const unsigned int c = c16 >> 4;
            if (c16 != 0xff0u) {
                double v;
                int ob;
                if (DM == DM_CONST) {
                    const DictEnt& e = *reinterpret_cast<const DictEnt*>(reinterpret_cast<const char*>(P.e) + c16);
                    v = e.v;
                    ob = e.offb;
                } else {
                    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"((c << 3) | sbase));
                    asm volatile("ld.shared.s32 %0, [%1+2048];" : "=r"(ob) : "r"((c << 2) | sbase));
                }
                sum[i] = fma(v, *reinterpret_cast<const double*>(xr[i] + ob), sum[i]);
            }
        }
    }
}

// PERSIST: an occupancy-sized grid, every warp strides over work items of NS = 2 adjacent slices, and the dependent chain
// slice_ptr -> codes -> x is software pipelined (the slice pointers of item i+2 and the code words of item i+1 are in flight
// while the x gathers of item i are issued).  !PERSIST: one work item per warp.  All indices are 32-bit (n_loc < 2^31 - 64;
// slice_ptr carries one padding entry so that slice_ptr[s+2] is always readable).
// one sweep y = A*x (+ Newton epilogue) over slices [slice_lo, slice_hi) by the whole grid; `sbase` from dict_stage
template <bool NEWTON, bool PERSIST, int DM>
__device__ __forceinline__ void selld_sweep(const int32_t* __restrict__ slice_ptr, const uint2* __restrict__ codes, const DictParam& P,
                                            uint32_t sbase, const double* x, const double* xprev, double* y, int slice_lo, int slice_hi,
                                            int n_loc, double shift, double pair) {
    constexpr int NS = kDictSlicesPerWarp;
    static_assert(NS == 2, "the pointer loads below fetch slice_ptr[s .. s+2]");
    const int lane = threadIdx.x & 31;
    const int items = (slice_hi - slice_lo + NS - 1) / NS;
    const int stride = PERSIST ? (int)gridDim.x * (kSpmvThreads / 32) : items;
    int it = (int)blockIdx.x * (kSpmvThreads / 32) + (threadIdx.x >> 5);
    const uint2* __restrict__ cl = codes + lane;

    auto load_ptrs = [&](int item, int32_t (&p)[NS], int32_t (&n)[NS]) {
        p[0] = p[1] = 0;
        n[0] = n[1] = 0;
        if (item < items) {
            const int s = slice_lo + item * NS;
            const int32_t a = __ldg(slice_ptr + s), b = __ldg(slice_ptr + s + 1), c = __ldg(slice_ptr + s + 2);
            p[0] = a;
            n[0] = b - a;
            p[1] = b;
            n[1] = (s + 1 < slice_hi) ? c - b : 0;
        }
    };
    auto load_codes = [&](const int32_t (&p)[NS], const int32_t (&n)[NS], int b, uint2 (&w)[NS]) {
#pragma unroll
        for (int i = 0; i < NS; ++i) w[i] = (b < n[i]) ? __ldg(cl + ((size_t)(p[i] + b) << 5)) : make_uint2(~0u, ~0u);
    };

    // prologue: pointers of items it and it+stride, code words of item it
    int32_t p0[NS], nb[NS], p0n[NS], nbn[NS];
    uint2 w[NS];
    load_ptrs(it, p0, nb);
    if (PERSIST) load_ptrs(it + stride, p0n, nbn);
    load_codes(p0, nb, 0, w);
    for (; it < items; it += stride) {
        uint2 wn[NS];
        int32_t p0nn[NS], nbnn[NS];
        if (PERSIST) {   // prefetch: code words of the next item, slice pointers of the one after
            load_codes(p0n, nbn, 0, wn);
            load_ptrs(it + 2 * stride, p0nn, nbnn);
        }
        const int row0 = (slice_lo + it * NS) * 32 + lane;
        double sum[NS];
        const char* xr[NS];
#pragma unroll
        for (int i = 0; i < NS; ++i) {
            sum[i] = 0.0;
            xr[i] = reinterpret_cast<const char*>(x + (row0 + 32 * i));
        }
        dict_fma8<DM, NS>(P, sbase, w, xr, sum);
        const int maxb = max(nb[0], nb[1]);
        for (int b = 1; b < maxb; ++b) {          // rows with more than 8 non-zeros
            load_codes(p0, nb, b, w);
            dict_fma8<DM, NS>(P, sbase, w, xr, sum);
        }
#pragma unroll
        for (int i = 0; i < NS; ++i) {
            const int r = row0 + 32 * i;
            if (r < n_loc && (i == 0 || slice_lo + it * NS + i < slice_hi)) {
                double v = sum[i];
                if (NEWTON) v = newton_epilogue(v, x[r], pair != 0.0 ? xprev[r] : 0.0, shift, pair);
                y[r] = v;
            }
        }
        if (PERSIST) {
#pragma unroll
            for (int i = 0; i < NS; ++i) {
                w[i] = wn[i];
                p0[i] = p0n[i]; nb[i] = nbn[i];
                p0n[i] = p0nn[i]; nbn[i] = nbnn[i];
            }
        }
    }
}

template <bool NEWTON, bool PERSIST, int DM>
__global__ void __launch_bounds__(kSpmvThreads, 5)
k_spmv_selld(const int32_t* __restrict__ slice_ptr, const uint2* __restrict__ codes, const __grid_constant__ DictParam P,
             const double* __restrict__ x, const double* __restrict__ xprev, double* __restrict__ y, int slice_lo,
             int slice_hi, int n_loc, double shift, double pair) {
    __shared__ __align__(4096) unsigned char sraw[DM == DM_CONST ? 16 : 4096];
    const uint32_t sbase = dict_stage<DM>(sraw, P);
    selld_sweep<NEWTON, PERSIST, DM>(slice_ptr, codes, P, sbase, x, xprev, y, slice_lo, slice_hi, n_loc, shift, pair);
}

// ---- slice patterns: a slice whose 32 rows hold the same (offset, value) entries needs no code bytes at all.  The interior of a
//      stencil is ONE pattern (7 entries, all lanes): its offsets and values live in registers for the whole kernel and a row costs
//      one 64-bit address add, one coalesced load and one DFMA per non-zero (3 instructions instead of 9, and 16 bytes of HBM
//      traffic per row instead of 24).  Other patterns (rows at the x-boundaries miss a neighbour on one lane: lane masks) are
//      read from the constant bank; slices without a pattern take the coded path.  Same products, same order: bit-identical.
struct __align__(16) PatEnt {
    double v;
    int offb;
    unsigned mask;
};
struct PatParam {                          // = calz_mat::h_pat
    PatEnt e0[8];                          // entries of pattern 0 (ascending offset)
    unsigned mask[32][8];                  // pattern p: lanes that hold entry k
};

// NS = 2 consecutive slices per warp and iteration (14 independent gathers in flight for a 7-point stencil); items are
// grid-strided: the whole grid sweeps one narrow window of rows, so the +-plane gathers of a stencil are L2 hits (measured
// alternative: waves in which every CTA owns 64 consecutive items -- better L1 reuse of the +-line gathers, but 0.83 vs 0.73 ms
// per C3 MPK).  CNT = number of
// entries of pattern 0 (compile time: the unrolled loops carry no count tests; values and byte offsets are constant-bank operands
// of the DFMA / address instructions).  The host only selects this kernel when pattern 0 holds every entry on every lane.
template <bool NEWTON, int CNT>
__global__ void __launch_bounds__(kSpmvThreads, 5)
k_spmv_selp(const uint8_t* __restrict__ spat, const int32_t* __restrict__ slice_ptr, const uint2* __restrict__ codes,
            const __grid_constant__ DictParam D, const __grid_constant__ PatParam PP, const double* __restrict__ x,
            const double* __restrict__ xprev, double* __restrict__ y, int slice_lo, int slice_hi, int n_loc, double shift, double pair,
            int prefetch, int phase) {
    const int lane = threadIdx.x & 31;
    const unsigned lanebit = 1u << lane;
    // item j (absolute) = slices 2j - phase and 2j - phase + 1.  phase (0/1, chosen per matrix at set-up) is the pairing that puts
    // most slices with a boundary pattern TOGETHER -- a 256-row grid line is 8 slices whose first and last one touch the boundary:
    // paired (7 | 0 of the next line) three items of four take the interior fast path, paired (0 1) ... (6 7) only two of four do.
    const int j0 = (slice_lo + phase) >> 1, items = ((slice_hi - 1 + phase) >> 1) - j0 + 1;
    const int stride = (int)gridDim.x * (kSpmvThreads / 32);
    const int wid = (int)blockIdx.x * (kSpmvThreads / 32) + (threadIdx.x >> 5);
    // Round k of the grid sweep covers items [k*stride, (k+1)*stride); inside a round the warps ROTATE by one item per round
    // (warp w takes item k*stride + (w + k) mod stride).  stride is a multiple of 8 and so is the period of the boundary slices of
    // a power-of-two grid line: without the rotation a warp would meet the same kind of item (all interior, or all boundary) in
    // every round and the boundary warps would finish last.
    auto item_of = [&](int k) -> int {
        int r = wid + k;
        if (r >= stride) r -= stride;
        return k * stride + r;
    };
    int k = 0;
    int it = item_of(0);
    auto pids_of = [&](int item) -> int {                 // both pattern numbers of an item in one 16-bit word
        if (item >= items) return 0xffff;
        const int sl = 2 * (j0 + item) - phase;
        const int a = (sl >= slice_lo) ? (int)__ldg(spat + sl) : 254;             // 254: no such slice in this launch
        const int b = (sl + 1 < slice_hi) ? (int)__ldg(spat + sl + 1) : 254;
        return a | (b << 8);
    };
    int pids = pids_of(it);
    const int pf_items = prefetch * stride;
    const long long pf_off = (long long)pf_items * 512 + PP.e0[CNT - 1].offb;
    for (; k * stride < items; ++k) {
        const int it_next = item_of(k + 1);
        const int pids_next = pids_of(it_next);
        if (it >= items) { it = it_next; pids = pids_next; continue; }      // only in the last, partial round
        const int sl = 2 * (j0 + it) - phase;
        const int row = sl * 32 + lane;
        const char* xr = reinterpret_cast<const char*>(x + row);
        // Of the CNT gathers of a row only the one with the largest offset is new data (a stencil re-reads everything else from
        // L1/L2), so the gathers in flight cover few DRAM bytes: pull the leading edge of x into L2 kPrefetch iterations ahead.
        if (prefetch > 0 && it + pf_items < items) {
            const char* pf = xr + pf_off;
            asm volatile("prefetch.global.L2 [%0];" ::"l"(pf));
            asm volatile("prefetch.global.L2 [%0];" ::"l"(pf + 256));
        }
        double sum[2] = {0.0, 0.0};
        if (pids == 0) {
            // two interior slices: every entry on every lane
            double xa[CNT], xb[CNT];
#pragma unroll
            for (int k = 0; k < CNT; ++k) {
                xa[k] = *reinterpret_cast<const double*>(xr + PP.e0[k].offb);
                xb[k] = *reinterpret_cast<const double*>(xr + 256 + PP.e0[k].offb);
            }
#pragma unroll
            for (int k = 0; k < CNT; ++k) {
                sum[0] = fma(PP.e0[k].v, xa[k], sum[0]);
                sum[1] = fma(PP.e0[k].v, xb[k], sum[1]);
            }
        } else {
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const int pid = (pids >> (8 * i)) & 0xff;
                const char* xi = xr + 256 * i;
                if (pid == 0) {                                       // interior slice next to a boundary one: no masks
                    double xa[CNT];
#pragma unroll
                    for (int k = 0; k < CNT; ++k) xa[k] = *reinterpret_cast<const double*>(xi + PP.e0[k].offb);
#pragma unroll
                    for (int k = 0; k < CNT; ++k) sum[i] = fma(PP.e0[k].v, xa[k], sum[i]);
                } else if (pid < 32) {
#pragma unroll
                    for (int k = 0; k < CNT; ++k)
                        if (PP.mask[pid][k] & lanebit) sum[i] = fma(PP.e0[k].v, *reinterpret_cast<const double*>(xi + PP.e0[k].offb), sum[i]);
                } else if (pid == 255) {
                    const int32_t p0 = __ldg(slice_ptr + sl + i), nb = __ldg(slice_ptr + sl + i + 1) - p0;
                    for (int b = 0; b < nb; ++b) {
                        const uint2 w = __ldg(codes + (((size_t)(p0 + b)) << 5) + lane);
#pragma unroll
                        for (int q = 0; q < 8; ++q) {
                            const unsigned int half = q < 4 ? w.x : w.y;
                            const unsigned int c16 = ((q & 3) == 0 ? (half << 4) : (half >> (8 * (q & 3) - 4))) & 0xff0u;
                            if (c16 != 0xff0u) {
                                const DictEnt& e = *reinterpret_cast<const DictEnt*>(reinterpret_cast<const char*>(D.e) + c16);
                                sum[i] = fma(e.v, *reinterpret_cast<const double*>(xi + e.offb), sum[i]);
                            }
                        }
                    }
                }
            }
        }
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const int r = row + 32 * i;
            if (r < n_loc && ((pids >> (8 * i)) & 0xff) != 254) {
                double v = sum[i];
                if (NEWTON) v = newton_epilogue(v, x[r], pair != 0.0 ? xprev[r] : 0.0, shift, pair);
                y[r] = v;
            }
        }
        pids = pids_next;
        it = it_next;
    }
}

template <bool NEWTON, int CNT>
int launch_selp_t(calz_mat* m, const double* x, const double* xp, double* y, int64_t s0, int64_t s1, double shift, double pair) {
    if (s1 <= s0) return CALZ_OK;
    calz_ctx* ctx = m->ctx;
    static int occ = 0;
    if (!occ) {
        CALZ_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_spmv_selp<NEWTON, CNT>, kSpmvThreads, 0));
        if (occ < 1) occ = 1;
    }
    const int64_t per_cta = 2 * (kSpmvThreads / 32);
    const unsigned grid = (unsigned)std::min<int64_t>((s1 - s0 + 1 + per_cta - 1) / per_cta, (int64_t)ctx->num_sms * occ);
    static_assert(sizeof(PatParam) == sizeof(m->h_pat), "pattern parameter block");
    k_spmv_selp<NEWTON, CNT><<<grid, kSpmvThreads, 0, ctx->stream>>>(m->d_slice_pat, m->d_slice_ptr, (const uint2*)m->d_codes,
                                                                      *(const DictParam*)m->h_dict, *(const PatParam*)m->h_pat, x, xp, y,
                                                                      (int)s0, (int)s1, (int)m->n_loc, shift, pair, (int)ctx->opt_mpk_prefetch,
                                                                      ctx->opt_mpk_pair_phase < 0 ? m->pat_phase : (int)(ctx->opt_mpk_pair_phase & 1));
    CALZ_LAUNCH_CHECK(ctx);
    return CALZ_OK;
}

template <bool NEWTON>
int launch_selp(calz_mat* m, const double* x, const double* xp, double* y, int64_t s0, int64_t s1, double shift, double pair) {
    switch (m->pat_cnt0) {
        case 1: return launch_selp_t<NEWTON, 1>(m, x, xp, y, s0, s1, shift, pair);
        case 2: return launch_selp_t<NEWTON, 2>(m, x, xp, y, s0, s1, shift, pair);
        case 3: return launch_selp_t<NEWTON, 3>(m, x, xp, y, s0, s1, shift, pair);
        case 4: return launch_selp_t<NEWTON, 4>(m, x, xp, y, s0, s1, shift, pair);
        case 5: return launch_selp_t<NEWTON, 5>(m, x, xp, y, s0, s1, shift, pair);
        case 6: return launch_selp_t<NEWTON, 6>(m, x, xp, y, s0, s1, shift, pair);
        case 7: return launch_selp_t<NEWTON, 7>(m, x, xp, y, s0, s1, shift, pair);
        default: return launch_selp_t<NEWTON, 8>(m, x, xp, y, s0, s1, shift, pair);
    }
}

// ---- the matrix powers kernel proper: ALL steps of one exchange group in ONE cooperative launch.  Step k sweeps the slices
//      [lo_k, hi_k) of column k-1 -> column k of the basis workspace; a grid-wide barrier (one atomic per CTA on a monotonically
//      increasing 64-bit counter, release/acquire fences) separates the steps.  No column is read before it is written inside the
//      launch, so the L1-cached x gathers of the single-step kernel stay valid.  Same arithmetic, same order: bit-identical.
constexpr int kMpkMaxSteps = 32;
struct MpkSteps {
    int nsteps, col0;                     // step i reads column col0 + i, writes column col0 + i + 1
    int lo[kMpkMaxSteps], hi[kMpkMaxSteps];
    double shift[kMpkMaxSteps], pair[kMpkMaxSteps];
};

__device__ __forceinline__ void grid_barrier(unsigned long long* counter, unsigned long long target) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(counter, 1ull);
        while (*(volatile unsigned long long*)counter < target) { }
        __threadfence();
    }
    __syncthreads();
}

template <bool NEWTON, int DM>
__global__ void __launch_bounds__(kSpmvThreads, 5)
k_mpk_selld(const int32_t* __restrict__ slice_ptr, const uint2* __restrict__ codes, const __grid_constant__ DictParam P,
            double* W, long long ldW, const __grid_constant__ MpkSteps st, int n_loc, unsigned long long* counter,
            unsigned long long base) {
    __shared__ __align__(4096) unsigned char sraw[DM == DM_CONST ? 16 : 4096];
    const uint32_t sbase = dict_stage<DM>(sraw, P);
    for (int i = 0; i < st.nsteps; ++i) {
        const int col = st.col0 + i;
        const double* x = W + (long long)col * ldW;
        const double* xp = col >= 1 ? W + (long long)(col - 1) * ldW : x;
        selld_sweep<NEWTON, true, DM>(slice_ptr, codes, P, sbase, x, xp, W + (long long)(col + 1) * ldW, st.lo[i], st.hi[i], n_loc,
                                      st.shift[i], st.pair[i]);
        if (i + 1 < st.nsteps) grid_barrier(counter, base + (unsigned long long)gridDim.x * (i + 1));
    }
}

// ---- dictionary-coded SELL with the x vector staged by TMA: a CTA owns xs_rows consecutive rows; one elected thread issues
//      one cp.async.bulk per merged x segment (own range + stencil halos / ghost zone) into shared memory, everybody waits on
//      the mbarrier, and the inner loop is two shared-memory reads and one FMA per non-zero -- no global gathers at all.
struct XsPlan {
    int rows, groups, total;
    int omin[8], len[8], base[8];
};

__device__ __forceinline__ uint32_t mpk_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <bool NEWTON>
__global__ void __launch_bounds__(kSpmvThreads)
k_spmv_selld_tma(const int32_t* __restrict__ slice_ptr, const uint2* __restrict__ codes, const double2* __restrict__ dict, int dict_size,
                 const int* __restrict__ xs_off, XsPlan plan, const double* __restrict__ x, const double* __restrict__ xprev,
                 double* __restrict__ y, int64_t slice_lo, int64_t slice_hi, int64_t n_loc, int64_t ldw, double shift, double pair) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double* xs = reinterpret_cast<double*>(smem_raw);
    double* sval = xs + plan.total;
    int* soff = reinterpret_cast<int*>(sval + 256);
    uint64_t* bar = reinterpret_cast<uint64_t*>(soff + 256);
    const int slices_per_cta = plan.rows / 32;
    const int64_t s_first = slice_lo + (int64_t)blockIdx.x * slices_per_cta;
    const int64_t r0 = s_first * 32;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mpk_smem_u32(bar)), "r"(1) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = threadIdx.x; i < 256; i += kSpmvThreads) {
        sval[i] = i < dict_size ? __ldg(dict + i).x : 0.0;
        soff[i] = i < dict_size ? __ldg(xs_off + i) : 0;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t bytes = 0;
        long long lo[8], hi[8];
        for (int g = 0; g < plan.groups; ++g) {
            lo[g] = max((long long)(r0 + plan.omin[g]), 0LL);
            hi[g] = min((long long)(r0 + plan.omin[g] + plan.len[g]), (long long)ldw);
            if (hi[g] > lo[g]) bytes += (uint32_t)((hi[g] - lo[g]) * 8);
        }
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mpk_smem_u32(bar)), "r"(bytes) : "memory");
        for (int g = 0; g < plan.groups; ++g) {
            if (hi[g] <= lo[g]) continue;
            double* dst = xs + plan.base[g] + (lo[g] - (r0 + plan.omin[g]));
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(mpk_smem_u32(dst)), "l"(x + lo[g]), "r"((uint32_t)((hi[g] - lo[g]) * 8)), "r"(mpk_smem_u32(bar)) : "memory");
        }
    }
    // each warp owns a run of consecutive slices, processed four at a time with all index/code loads issued up front;
    // the first batch is fetched BEFORE waiting for the x segments so that the two latencies overlap
    constexpr int NS = 4;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int per_warp = slices_per_cta / (kSpmvThreads / 32);
    bool waited = false;
    for (int i0 = warp * per_warp; i0 < (warp + 1) * per_warp; i0 += NS) {
        int32_t p0[NS], nb[NS];
        int maxb = 0;
#pragma unroll
        for (int i = 0; i < NS; ++i) {
            const int64_t slice = s_first + i0 + i;
            const bool ok = i0 + i < (warp + 1) * per_warp && slice < slice_hi;
            p0[i] = ok ? __ldg(slice_ptr + slice) : 0;
            nb[i] = ok ? __ldg(slice_ptr + slice + 1) - p0[i] : 0;
            maxb = max(maxb, nb[i]);
        }
        double sum[NS];
#pragma unroll
        for (int i = 0; i < NS; ++i) sum[i] = 0.0;
        for (int b = 0; b < maxb; ++b) {
            uint2 w[NS];
#pragma unroll
            for (int i = 0; i < NS; ++i) w[i] = (b < nb[i]) ? __ldg(codes + (int64_t)(p0[i] + b) * 32 + lane) : make_uint2(~0u, ~0u);
            if (!waited) {   // everybody waits for the segments (once)
                uint32_t done;
                do {
                    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                                 : "=r"(done) : "r"(mpk_smem_u32(bar)), "r"(0) : "memory");
                } while (!done);
                waited = true;
            }
#pragma unroll
            for (int q = 0; q < 8; ++q) {
#pragma unroll
                for (int i = 0; i < NS; ++i) {
                    const unsigned int half = q < 4 ? w[i].x : w[i].y;
                    const unsigned int c = (half >> (8 * (q & 3))) & 0xffu;
                    if (c != 255u) sum[i] = fma(sval[c], xs[soff[c] + (i0 + i) * 32 + lane], sum[i]);
                }
            }
        }
#pragma unroll
        for (int i = 0; i < NS; ++i) {
            const int64_t r = r0 + (int64_t)(i0 + i) * 32 + lane;
            if (nb[i] >= 0 && i0 + i < (warp + 1) * per_warp && s_first + i0 + i < slice_hi && r < n_loc) {
                double v = sum[i];
                if (NEWTON) v = newton_epilogue(v, x[r], pair != 0.0 ? xprev[r] : 0.0, shift, pair);
                y[r] = v;
            }
        }
    }
}

// ---- CSR: L lanes per row (L chosen from the mean row length), shuffle reduction
template <int L, bool NEWTON>
__global__ void __launch_bounds__(kSpmvThreads)
k_spmv_csr(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col, const double* __restrict__ val,
           const double* __restrict__ x, const double* __restrict__ xprev, double* __restrict__ y,
           int64_t row_lo, int64_t row_hi, double shift, double pair) {
    const int64_t gid = (int64_t)blockIdx.x * kSpmvThreads + threadIdx.x;
    const int64_t row = row_lo + gid / L;
    const int sub = (int)(gid % L);
    double sum = 0.0;
    if (row < row_hi) {
        const int32_t e0 = __ldg(rowptr + row), e1 = __ldg(rowptr + row + 1);
        for (int32_t e = e0 + sub; e < e1; e += L) sum = fma(__ldg(val + e), x[__ldg(col + e)], sum);
    }
#pragma unroll
    for (int o = L / 2; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    if (row < row_hi && sub == 0) {
        if (NEWTON) sum = newton_epilogue(sum, x[row], pair != 0.0 ? xprev[row] : 0.0, shift, pair);
        y[row] = sum;
    }
}

// ---- long rows (hubs of a power-law graph): one CTA per segment of <= 4096 entries, fixed reduction order; then one thread per
//      long row adds its segments in order and applies the epilogue.  Runs after the main kernel (which saw these rows as empty).
__global__ void __launch_bounds__(256)
k_spmv_long_seg(const int32_t* __restrict__ segptr, const int32_t* __restrict__ segrow, const int32_t* __restrict__ lrow,
                const int32_t* __restrict__ col, const double* __restrict__ val, const double* __restrict__ x,
                double* __restrict__ part, int lo, int hi) {
    const int g = blockIdx.x;
    const int row = __ldg(lrow + __ldg(segrow + g));
    if (row < lo || row >= hi) return;
    const int e0 = __ldg(segptr + g), e1 = __ldg(segptr + g + 1);
    double sum = 0.0;
    for (int e = e0 + threadIdx.x; e < e1; e += 256) sum = fma(__ldg(val + e), x[__ldg(col + e)], sum);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    __shared__ double ws[8];
    if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = sum;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
#pragma unroll
        for (int w = 0; w < 8; ++w) t += ws[w];
        part[g] = t;
    }
}

template <bool NEWTON>
__global__ void k_spmv_long_fin(int n_long, const int32_t* __restrict__ seg0, const int32_t* __restrict__ lrow,
                                const double* __restrict__ part, const double* __restrict__ x, const double* __restrict__ xprev,
                                double* __restrict__ y, int lo, int hi, double shift, double pair) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_long) return;
    const int row = __ldg(lrow + r);
    if (row < lo || row >= hi) return;
    double sum = 0.0;
    for (int g = __ldg(seg0 + r); g < __ldg(seg0 + r + 1); ++g) sum += part[g];
    if (NEWTON) sum = newton_epilogue(sum, x[row], pair != 0.0 ? xprev[row] : 0.0, shift, pair);
    y[row] = sum;
}

__global__ void k_pack(const double* __restrict__ x, const int32_t* __restrict__ idx, double* __restrict__ buf, int64_t n) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) buf[i] = x[idx[i]];
}

template <bool NEWTON>
int launch_csr(calz_mat* m, const double* x, const double* xp, double* y, int64_t lo, int64_t hi, double shift, double pair) {
    calz_ctx* ctx = m->ctx;
    const int L = m->csr_lanes;
    const int64_t threads = (hi - lo) * L;
    const unsigned grid = (unsigned)((threads + kSpmvThreads - 1) / kSpmvThreads);
#define CALZ_CSR_CASE(LL)                                                                                     \
    case LL:                                                                                                  \
        k_spmv_csr<LL, NEWTON><<<grid, kSpmvThreads, 0, ctx->stream>>>(m->d_rowptr, m->d_colind, m->d_val, x, \
                                                                       xp, y, lo, hi, shift, pair);           \
        break;
    switch (L) {
        CALZ_CSR_CASE(1) CALZ_CSR_CASE(2) CALZ_CSR_CASE(4) CALZ_CSR_CASE(8) CALZ_CSR_CASE(16) CALZ_CSR_CASE(32)
        default: return set_error(ctx, CALZ_ERR_BADARG, "csr_lanes must be a power of two <= 32");
    }
#undef CALZ_CSR_CASE
    CALZ_LAUNCH_CHECK(ctx);
    return CALZ_OK;
}

template <bool NEWTON, bool PERSIST, int DM>
int launch_selld_t(calz_mat* m, const double* x, const double* xp, double* y, int64_t s0, int64_t s1, double shift, double pair) {
    if (s1 <= s0) return CALZ_OK;
    calz_ctx* ctx = m->ctx;
    const int64_t per_cta = (int64_t)(kSpmvThreads / 32) * kDictSlicesPerWarp;
    unsigned grid = (unsigned)((s1 - s0 + per_cta - 1) / per_cta);
    if (PERSIST) {
        static int occ = 0;
        if (!occ) {
            CALZ_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_spmv_selld<NEWTON, PERSIST, DM>, kSpmvThreads, 0));
            if (occ < 1) occ = 1;
        }
        const unsigned cap = (unsigned)(ctx->num_sms * occ);
        if (grid > cap) grid = cap;
    }
    k_spmv_selld<NEWTON, PERSIST, DM><<<grid, kSpmvThreads, 0, ctx->stream>>>(m->d_slice_ptr, (const uint2*)m->d_codes, *(const DictParam*)m->h_dict,
                                                                             x, xp, y, (int)s0, (int)s1, (int)m->n_loc, shift, pair);
    CALZ_LAUNCH_CHECK(ctx);
    return CALZ_OK;
}

int launch_selld(calz_mat* m, const double* x, const double* xp, double* y, int64_t lo, int64_t hi, bool newton, double shift, double pair) {
    calz_ctx* ctx = m->ctx;
    static_assert(sizeof(DictParam) == sizeof(m->h_dict), "dictionary parameter block");
    const int64_t s0 = lo / 32, s1 = (hi + 31) / 32;
    int dm = (int)ctx->opt_mpk_dict_mode;
    if (dm < 0) dm = m->dict_uniform >= 0.75 ? DM_CONST : DM_SHARED;
    const int key = (newton ? 1 : 0) | (ctx->opt_mpk_persist ? 2 : 0) | (dm << 2);
    if (!newton) { shift = 0.0; pair = 0.0; }
    if (ctx->opt_mpk_patterns && dm == DM_CONST && m->d_slice_pat && m->pat_cover >= 0.5 && m->pat_cnt0 >= 1) {   // most slices have a pattern
        return newton ? launch_selp<true>(m, x, xp, y, s0, s1, shift, pair) : launch_selp<false>(m, x, xp, y, s0, s1, shift, pair);
    }
#define CALZ_SELLD_CASE(NW, PS, DM) \
    case ((NW) | ((PS) << 1) | ((DM) << 2)): return launch_selld_t<NW != 0, PS != 0, DM>(m, x, xp, y, s0, s1, shift, pair);
    switch (key) {
        CALZ_SELLD_CASE(0, 0, 0) CALZ_SELLD_CASE(1, 0, 0) CALZ_SELLD_CASE(0, 1, 0) CALZ_SELLD_CASE(1, 1, 0)
        CALZ_SELLD_CASE(0, 0, 2) CALZ_SELLD_CASE(1, 0, 2) CALZ_SELLD_CASE(0, 1, 2) CALZ_SELLD_CASE(1, 1, 2)
        default: return set_error(ctx, CALZ_ERR_BADARG, "mpk_dict_mode must be -1 (auto), 0 (shared) or 2 (constant bank)");
    }
#undef CALZ_SELLD_CASE
}

int spmv_main(calz_mat* m, const double* x, const double* xp, double* y, int64_t lo, int64_t hi, bool newton, double shift, double pair);

// steps k0 .. k0+ns-1 (1-based) of the MPK in one cooperative launch (dictionary layout, persistent grid)
template <bool NEWTON, int DM>
int launch_mpk_selld_t(calz_mat* m, const MpkSteps& st) {
    calz_ctx* ctx = m->ctx;
    static int occ = 0;
    if (!occ) {
        CALZ_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_mpk_selld<NEWTON, DM>, kSpmvThreads, 0));
        if (occ < 1) occ = 1;
    }
    int maxs = 0;
    for (int i = 0; i < st.nsteps; ++i) maxs = std::max(maxs, st.hi[i] - st.lo[i]);
    const int64_t per_cta = (int64_t)(kSpmvThreads / 32) * kDictSlicesPerWarp;
    unsigned grid = (unsigned)std::max<int64_t>(1, std::min<int64_t>((maxs + per_cta - 1) / per_cta, (int64_t)ctx->num_sms * occ));
    if (!m->d_gridbar) {
        CALZ_CUDA(ctx, cudaMalloc(&m->d_gridbar, 64));
        CALZ_CUDA(ctx, cudaMemsetAsync(m->d_gridbar, 0, 64, ctx->stream));
        m->gridbar_base = 0;
    }
    const int32_t* sp = m->d_slice_ptr;
    const uint2* codes = (const uint2*)m->d_codes;
    DictParam* P = (DictParam*)m->h_dict;
    double* W = m->d_W;
    long long ld = m->ldW;
    int n_loc = (int)m->n_loc;
    unsigned long long* ctr = m->d_gridbar;
    unsigned long long base = m->gridbar_base;
    MpkSteps stc = st;
    void* args[] = {(void*)&sp, (void*)&codes, (void*)P, (void*)&W, (void*)&ld, (void*)&stc, (void*)&n_loc, (void*)&ctr, (void*)&base};
    CALZ_CUDA(ctx, cudaLaunchCooperativeKernel((const void*)k_mpk_selld<NEWTON, DM>, dim3(grid), dim3(kSpmvThreads), args, 0, ctx->stream));
    ctx->launches++;
    m->gridbar_base += (unsigned long long)grid * (unsigned long long)(st.nsteps - 1);
    return CALZ_OK;
}

int launch_mpk_selld(calz_mat* m, const MpkSteps& st, bool newton) {
    calz_ctx* ctx = m->ctx;
    int dm = (int)ctx->opt_mpk_dict_mode;
    if (dm < 0) dm = m->dict_uniform >= 0.75 ? DM_CONST : DM_SHARED;
    if (dm == DM_CONST) return newton ? launch_mpk_selld_t<true, DM_CONST>(m, st) : launch_mpk_selld_t<false, DM_CONST>(m, st);
    return newton ? launch_mpk_selld_t<true, DM_SHARED>(m, st) : launch_mpk_selld_t<false, DM_SHARED>(m, st);
}

// one SpMV step on local rows [lo,hi) (already aligned to the layout granule)
int spmv_step(calz_mat* m, const double* x, const double* xp, double* y, int64_t lo, int64_t hi, bool newton,
              double shift, double pair) {
    if (hi <= lo) return CALZ_OK;
    CALZ_TRY(spmv_main(m, x, xp, y, lo, hi, newton, shift, pair));
    if (m->n_long) {
        calz_ctx* ctx = m->ctx;
        k_spmv_long_seg<<<(unsigned)m->n_long_seg, 256, 0, ctx->stream>>>(m->d_long_segptr, m->d_long_segrow, m->d_long_row, m->d_long_col,
                                                                          m->d_long_val, x, m->d_long_part, (int)lo, (int)hi);
        CALZ_LAUNCH_CHECK(ctx);
        const unsigned g = (unsigned)((m->n_long + 127) / 128);
        if (newton)
            k_spmv_long_fin<true><<<g, 128, 0, ctx->stream>>>((int)m->n_long, m->d_long_seg0, m->d_long_row, m->d_long_part, x, xp, y,
                                                              (int)lo, (int)hi, shift, pair);
        else
            k_spmv_long_fin<false><<<g, 128, 0, ctx->stream>>>((int)m->n_long, m->d_long_seg0, m->d_long_row, m->d_long_part, x, xp, y,
                                                               (int)lo, (int)hi, 0.0, 0.0);
        CALZ_LAUNCH_CHECK(ctx);
    }
    return CALZ_OK;
}

int spmv_main(calz_mat* m, const double* x, const double* xp, double* y, int64_t lo, int64_t hi, bool newton,
              double shift, double pair) {
    calz_ctx* ctx = m->ctx;
    if (m->layout == CALZ_LAYOUT_SELL_DICT && m->xs_rows > 0 && ctx->opt_mpk_tma_x && m->W_pad == 0) {   // bulk copies need 16-B aligned x segments
        const int64_t s0 = lo / 32, s1 = (hi + 31) / 32;
        const int spc = m->xs_rows / 32;
        const unsigned grid = (unsigned)((s1 - s0 + spc - 1) / spc);
        XsPlan plan{};
        plan.rows = m->xs_rows; plan.groups = m->xs_groups; plan.total = m->xs_total;
        for (int g = 0; g < m->xs_groups; ++g) { plan.omin[g] = m->xs_omin[g]; plan.len[g] = m->xs_len[g]; plan.base[g] = m->xs_base[g]; }
        const size_t smem = (size_t)m->xs_total * 8 + 256 * 8 + 256 * 4 + 64;
        const uint2* codes = (const uint2*)m->d_codes;
        const double2* dict = (const double2*)m->d_dict;
        if (newton) {
            CALZ_CUDA(ctx, cudaFuncSetAttribute(k_spmv_selld_tma<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            k_spmv_selld_tma<true><<<grid, kSpmvThreads, smem, ctx->stream>>>(m->d_slice_ptr, codes, dict, m->dict_size, m->d_xs_off, plan,
                                                                              x, xp, y, s0, s1, m->n_loc, m->ldW, shift, pair);
        } else {
            CALZ_CUDA(ctx, cudaFuncSetAttribute(k_spmv_selld_tma<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            k_spmv_selld_tma<false><<<grid, kSpmvThreads, smem, ctx->stream>>>(m->d_slice_ptr, codes, dict, m->dict_size, m->d_xs_off, plan,
                                                                               x, xp, y, s0, s1, m->n_loc, m->ldW, 0.0, 0.0);
        }
        CALZ_LAUNCH_CHECK(ctx);
        return CALZ_OK;
    }
    if (m->layout == CALZ_LAYOUT_SELL_DICT) return launch_selld(m, x, xp, y, lo, hi, newton, shift, pair);
    if (m->layout == CALZ_LAYOUT_SELL) {
        const int64_t s0 = lo / 32, s1 = (hi + 31) / 32;
        const unsigned grid = (unsigned)((s1 - s0 + kSpmvThreads / 32 - 1) / (kSpmvThreads / 32));
        if (newton)
            k_spmv_sell<true><<<grid, kSpmvThreads, 0, ctx->stream>>>(m->d_slice_ptr, m->d_sell_col, m->d_sell_val, m->d_perm,
                                                                      x, xp, y, s0, s1, m->n_loc, shift, pair);
        else
            k_spmv_sell<false><<<grid, kSpmvThreads, 0, ctx->stream>>>(m->d_slice_ptr, m->d_sell_col, m->d_sell_val, m->d_perm,
                                                                       x, xp, y, s0, s1, m->n_loc, 0.0, 0.0);
        CALZ_LAUNCH_CHECK(ctx);
        return CALZ_OK;
    }
    return newton ? launch_csr<true>(m, x, xp, y, lo, hi, shift, pair) : launch_csr<false>(m, x, xp, y, lo, hi, 0.0, 0.0);
}

// one level-L halo exchange of workspace column `col` (L = s: the ONE exchange of an outer step, the communication-avoiding point)
int halo_exchange(calz_mat* m, double* W, int col) {
    calz_ctx* ctx = m->ctx;
    const int P = ctx->nranks;
    if (P <= 1) return CALZ_OK;
    if (m->p2p_halo) return p2p_halo_exchange(m, W, col);      // push into the peers' ghost zones over NVLink (p2p.cu)
    double* w = W + (int64_t)col * m->ldW;
    for (int q = 0; q < P; ++q)
        if (m->send_cnt[q] && !m->send_contig[q]) {
            const int64_t n = m->send_cnt[q];
            k_pack<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(w, m->d_send_idx + m->send_off[q],
                                                                        m->d_send_buf + m->send_off[q], n);
            CALZ_LAUNCH_CHECK(ctx);
        }
    CALZ_NCCL(ctx, ctx->nccl->GroupStart());
    for (int q = 0; q < P; ++q) {
        if (m->send_cnt[q]) {
            const double* src = m->send_contig[q] ? w + (m->own_off + (m->send_glob[q][0] - m->row_lo))
                                                  : m->d_send_buf + m->send_off[q];
            CALZ_NCCL(ctx, ctx->nccl->Send(src, (size_t)m->send_cnt[q], ncclFloat64, q, ctx->comm, ctx->stream));
        }
        if (m->recv_cnt[q])
            CALZ_NCCL(ctx, ctx->nccl->Recv(w + m->recv_off[q], (size_t)m->recv_cnt[q], ncclFloat64, q, ctx->comm, ctx->stream));
    }
    CALZ_NCCL(ctx, ctx->nccl->GroupEnd());
    return CALZ_OK;
}

struct Shifts {
    std::vector<double> re, pair;   // per step: real shift, im^2 factor (0 if none)
    bool newton = false;
};

int make_shifts(calz_ctx* ctx, int s, const double* re, const double* im, int modifiedp, int monomial, Shifts& out) {
    out.re.assign(s, 0.0);
    out.pair.assign(s, 0.0);
    out.newton = !monomial;
    if (monomial) return CALZ_OK;
    if (!re) return set_error(ctx, CALZ_ERR_BADARG, "mpk_newton: shift_re is NULL");
    for (int k = 0; k < s; ++k) {
        const double i = im ? im[k] : 0.0;
        out.re[k] = re[k];
        if (i != 0.0) {
            if (!modifiedp)
                return set_error(ctx, CALZ_ERR_UNSUPPORTED,
                                 "matrix_powers_newton with modifiedp=0 and a complex shift needs complex vectors");
            if (i < 0.0) {
                if (k == 0) return set_error(ctx, CALZ_ERR_SHIFT, "k==1, but shift %e has a negative imaginary part", re[k]);
                out.pair[k] = i * i;     // matrix_powers_newton.m:40-41
            }
        }
    }
    return CALZ_OK;
}

int mpk_run(calz_mat* m, const double* v, int s, const Shifts& sh) {
    calz_ctx* ctx = m->ctx;
    if (s < 1 || s > m->s_max) return set_error(ctx, CALZ_ERR_BADARG, "mpk: s=%d outside [1,%d] (s_max of the matrix)", s, m->s_max);
    CALZ_CUDA(ctx, cudaSetDevice(ctx->device));
    double* W = m->d_W;
    const int64_t ld = m->ldW;
    if (v != W + m->own_off)
        CALZ_CUDA(ctx, cudaMemcpyAsync(W + m->own_off, v, (size_t)m->n_own * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
    const int P = ctx->nranks;
    const int L = (P > 1) ? std::max(1, std::min(m->halo_level, s)) : s;      // steps per halo exchange

    const int64_t gran = (m->layout == CALZ_LAYOUT_SELL) ? (m->d_perm ? m->sell_sigma : 32) : (m->layout == CALZ_LAYOUT_SELL_DICT ? 32 : 1);
    // step k is the j-th of its exchange group: it has to produce the rows up to level min(L - j, s - k)
    auto lev_of = [&](int k) { return std::min(L - ((k - 1) % L + 1), s - k); };
    auto lo_of = [&](int k) { return (m->hull_lo[lev_of(k)] / gran) * gran; };
    auto hi_of = [&](int k) { return std::min<int64_t>(m->n_loc, round_up(m->hull_hi[lev_of(k)], gran)); };
    auto exchange_before = [&](int k) -> int {
        if ((k - 1) % L != 0) return CALZ_OK;
        // the conjugate-pair term of step k reads column k-2 on the ghost rows too, and the last step of the previous group
        // only produced its owned rows
        if (k >= 2 && sh.pair[k - 1] != 0.0) CALZ_TRY(halo_exchange(m, W, k - 2));
        return halo_exchange(m, W, k - 1);
    };
    auto ack_after = [&](int k) -> int {
        if (m->p2p_halo && (k - 1) % L == 0) return p2p_halo_ack(m);       // ghosts consumed: the owners may push the next ones
        return CALZ_OK;
    };
    auto step = [&](int k, int64_t lo, int64_t hi) -> int {
        lo = std::max(lo, lo_of(k));
        hi = std::min(hi, hi_of(k));
        const double* x = W + (int64_t)(k - 1) * ld;
        const double* xp = (k >= 2) ? W + (int64_t)(k - 2) * ld : x;
        return spmv_step(m, x, xp, W + (int64_t)k * ld, lo, hi, sh.newton, sh.re[k - 1], sh.pair[k - 1]);
    };

    // Temporal blocking through the 126 MB L2: process the rows in chunks whose slice of A stays resident,
    // skewing step k by (k-1) bandwidths so that every dependency is already computed.
    const int64_t bytes_per_row = m->n_loc ? (12 * m->nnz_loc) / m->n_loc + 24 : 0;
    const int64_t bwid = round_up(std::max<int64_t>(m->bandwidth, 1), gran);
    int64_t chunk_rows = 0;
    if (ctx->opt_l2_chunk_bytes > 0 && bytes_per_row > 0 && L >= s) {
        chunk_rows = round_up(std::max<int64_t>(ctx->opt_l2_chunk_bytes / bytes_per_row, gran), gran);
        if (chunk_rows < 2 * bwid || chunk_rows >= m->n_loc) chunk_rows = 0;   // band too wide / matrix fits: plain sweeps
    }
    const bool fused = chunk_rows == 0 && m->layout == CALZ_LAYOUT_SELL_DICT && ctx->opt_mpk_fused_steps && ctx->opt_mpk_persist &&
                       !(m->xs_rows > 0 && ctx->opt_mpk_tma_x && m->W_pad == 0) && m->n_long == 0 && s <= kMpkMaxSteps;
    if (fused) {
        // all steps of an exchange group in ONE cooperative launch (grid barrier between steps)
        for (int k0 = 1; k0 <= s; k0 += L) {
            const int ns = std::min(L, s - k0 + 1);
            CALZ_TRY(exchange_before(k0));
            MpkSteps st{};
            st.nsteps = ns;
            st.col0 = k0 - 1;
            for (int i = 0; i < ns; ++i) {
                const int k = k0 + i;
                st.lo[i] = (int)(lo_of(k) / 32);
                st.hi[i] = (int)((hi_of(k) + 31) / 32);
                st.shift[i] = sh.newton ? sh.re[k - 1] : 0.0;
                st.pair[i] = sh.newton ? sh.pair[k - 1] : 0.0;
            }
            CALZ_TRY(launch_mpk_selld(m, st, sh.newton));
            CALZ_TRY(ack_after(k0));
        }
    } else if (chunk_rows == 0) {
        for (int k = 1; k <= s; ++k) {
            CALZ_TRY(exchange_before(k));
            CALZ_TRY(step(k, 0, m->n_loc));
            CALZ_TRY(ack_after(k));
        }
    } else {
        CALZ_TRY(exchange_before(1));
        const int64_t span = m->n_loc + (int64_t)(s - 1) * bwid;
        for (int64_t c0 = 0; c0 < span; c0 += chunk_rows)
            for (int k = 1; k <= s; ++k) {
                const int64_t lo = c0 - (int64_t)(k - 1) * bwid, hi = lo + chunk_rows;
                if (hi <= 0 || lo >= m->n_loc) continue;
                CALZ_TRY(step(k, std::max<int64_t>(lo, 0), std::min<int64_t>(hi, m->n_loc)));
            }
        CALZ_TRY(ack_after(1));
    }
    return CALZ_OK;
}

int copy_out(calz_mat* m, int col0, int ncols, double* V, int64_t ldV) {
    calz_ctx* ctx = m->ctx;
    const double* src = m->d_W + m->own_off + (int64_t)col0 * m->ldW;
    if (V == src && ldV == m->ldW) return CALZ_OK;
    if (ldV < m->n_own) return set_error(ctx, CALZ_ERR_BADARG, "ldV < n_own");
    CALZ_CUDA(ctx, cudaMemcpy2DAsync(V, (size_t)ldV * sizeof(double), src, (size_t)m->ldW * sizeof(double),
                                     (size_t)m->n_own * sizeof(double), (size_t)ncols, cudaMemcpyDeviceToDevice, ctx->stream));
    return CALZ_OK;
}

}  // namespace

extern "C" {

int calz_mpk_inplace(calz_mat* m, const double* v, int s, const double* shift_re, const double* shift_im,
                     int modifiedp, int monomial, double** V, int64_t* ldV) {
    if (!m || !v) return set_error(m ? m->ctx : nullptr, CALZ_ERR_BADARG, "calz_mpk_inplace: bad arguments");
    Shifts sh;
    CALZ_TRY(make_shifts(m->ctx, s, shift_re, shift_im, modifiedp, monomial, sh));
    CALZ_TRY(mpk_run(m, v, s, sh));
    if (V) *V = m->d_W + m->own_off;
    if (ldV) *ldV = m->ldW;
    return CALZ_OK;
}

int calz_mpk_newton(calz_mat* m, const double* v, int s, const double* shift_re, const double* shift_im,
                    int modifiedp, double* V, int64_t ldV) {
    if (!m || !v || !V) return set_error(m ? m->ctx : nullptr, CALZ_ERR_BADARG, "calz_mpk_newton: bad arguments");
    CALZ_TRY(calz_mpk_inplace(m, v, s, shift_re, shift_im, modifiedp, 0, nullptr, nullptr));
    return copy_out(m, 0, s + 1, V, ldV);
}

int calz_mpk_monomial(calz_mat* m, const double* q, int s, double* V, int64_t ldV) {
    if (!m || !q || !V) return set_error(m ? m->ctx : nullptr, CALZ_ERR_BADARG, "calz_mpk_monomial: bad arguments");
    CALZ_TRY(calz_mpk_inplace(m, q, s, nullptr, nullptr, 0, 1, nullptr, nullptr));
    return copy_out(m, 1, s, V, ldV);       // matrix_powers_monomial.m:7 -- q itself is not returned
}

int calz_spmv(calz_mat* m, const double* x, double* y) {
    if (!m || !x || !y) return set_error(m ? m->ctx : nullptr, CALZ_ERR_BADARG, "calz_spmv: bad arguments");
    CALZ_TRY(calz_mpk_inplace(m, x, 1, nullptr, nullptr, 0, 1, nullptr, nullptr));
    return copy_out(m, 1, 1, y, m->n_own);
}

// ---- host-pointer flavours (what the MEX gateways call): H2D of the vector, D2H of the basis
static int mpk_host(calz_mat* m, const double* v, int s, const double* re, const double* im, int modifiedp,
                    int monomial, int col0, int ncols, double* V, int64_t ldV) {
    if (!m || !v || !V) return set_error(m ? m->ctx : nullptr, CALZ_ERR_BADARG, "mpk_host: bad arguments");
    calz_ctx* ctx = m->ctx;
    if (ldV < m->n_own) return set_error(ctx, CALZ_ERR_BADARG, "ldV < n");
    CALZ_CUDA(ctx, cudaSetDevice(ctx->device));
    double* w0 = m->d_W + m->own_off;
    CALZ_CUDA(ctx, cudaMemcpyAsync(w0, v, (size_t)m->n_own * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    CALZ_TRY(calz_mpk_inplace(m, w0, s, re, im, modifiedp, monomial, nullptr, nullptr));
    CALZ_CUDA(ctx, cudaMemcpy2DAsync(V, (size_t)ldV * sizeof(double), w0 + (int64_t)col0 * m->ldW,
                                     (size_t)m->ldW * sizeof(double), (size_t)m->n_own * sizeof(double), (size_t)ncols,
                                     cudaMemcpyDeviceToHost, ctx->stream));
    CALZ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return CALZ_OK;
}

int calz_spmv_host(calz_mat* m, const double* x, double* y) {
    return mpk_host(m, x, 1, nullptr, nullptr, 0, 1, 1, 1, y, m ? m->n_own : 0);
}

int calz_mpk_monomial_host(calz_mat* m, const double* q, int s, double* V, int64_t ldV) {
    return mpk_host(m, q, s, nullptr, nullptr, 0, 1, 1, s, V, ldV);
}

int calz_mpk_newton_host(calz_mat* m, const double* v, int s, const double* shift_re, const double* shift_im,
                         int modifiedp, double* V, int64_t ldV) {
    return mpk_host(m, v, s, shift_re, shift_im, modifiedp, 0, 0, s + 1, V, ldV);
}

}  // extern "C"

// EXPERIMENTAL dictionary-SELL SpMV with a cached code word (opt-in: `mpk_uniform_fast=1`).  Kept in its own translation unit
// so that the measured kernels of mpk.cu stay byte-identical; the few shared definitions are repeated here and pinned by
// static_asserts against matrix.h.
#include <algorithm>

#include "matrix.h"

using namespace calz;

namespace {

constexpr int kSpmvThreads = 256;
constexpr int kDictSlicesPerWarp = 2;

struct __align__(16) DictEnt {
    double v;
    int offb;                             // column offset pre-scaled to bytes
    int pad;
};
struct DictParam {                        // 4 KB, passed by value as a __grid_constant__ kernel parameter
    DictEnt e[256];
};

__device__ __forceinline__ double newton_epilogue(double w, double xi, double xp, double shift, double pair) {
    // matrix_powers_newton.m:34,40-41,43 -- two (three) separately rounded operations, no contraction (as in mpk.cu)
    double r = __dsub_rn(w, __dmul_rn(shift, xi));
    if (pair != 0.0) r = __dadd_rn(r, __dmul_rn(pair, xp));
    return r;
}

// ---- EXPERIMENTAL (opt-in, `mpk_uniform_fast=1`; not yet measured on hardware): the persistent constant-bank kernel with a
//      warp-level cache of ONE decoded code word.  In a stencil almost every slice carries the same 8 code bytes in all 32
//      lanes (the interior pattern), so the 8 dictionary entries of that word are kept in registers and a matching slice costs
//      a 64-bit add, an LDG and a DFMA per non-zero (4 instructions instead of 9).  Any other slice takes the generic path;
//      a slice whose lanes agree on a DIFFERENT word re-fills the cache.  Same products, same summation order: bit-identical.
template <bool NEWTON>
__global__ void __launch_bounds__(kSpmvThreads, 3)
k_spmv_selld_ufast(const int32_t* __restrict__ slice_ptr, const uint2* __restrict__ codes, const __grid_constant__ DictParam P,
                   const double* __restrict__ x, const double* __restrict__ xprev, double* __restrict__ y, int slice_lo,
                   int slice_hi, int n_loc, double shift, double pair) {
    constexpr int NS = kDictSlicesPerWarp;
    static_assert(NS == 2, "the pointer loads below fetch slice_ptr[s .. s+2]");
    const int lane = threadIdx.x & 31;
    const int items = (slice_hi - slice_lo + NS - 1) / NS;
    const int stride = (int)gridDim.x * (kSpmvThreads / 32);
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
    auto code_at = [](const uint2& w, int q) { return ((q < 4 ? w.x : w.y) >> (8 * (q & 3))) & 0xffu; };

    // the cached word: all padding to start with (matches nothing that carries work)
    uint2 cw = make_uint2(~0u, ~0u);
    double cv[8];
    int cob[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) { cv[q] = 0.0; cob[q] = 0; }
    int cnz = 0;                                  // leading non-padding codes of the cached word (padding is trailing)

    int32_t p0[NS], nb[NS], p0n[NS], nbn[NS];
    uint2 w[NS];
    load_ptrs(it, p0, nb);
    load_ptrs(it + stride, p0n, nbn);
    load_codes(p0, nb, 0, w);
    for (; it < items; it += stride) {
        uint2 wn[NS];
        int32_t p0nn[NS], nbnn[NS];
        load_codes(p0n, nbn, 0, wn);
        load_ptrs(it + 2 * stride, p0nn, nbnn);
        const int row0 = (slice_lo + it * NS) * 32 + lane;
        double sum[NS];
        const char* xr[NS];
#pragma unroll
        for (int i = 0; i < NS; ++i) {
            sum[i] = 0.0;
            xr[i] = reinterpret_cast<const char*>(x + (row0 + 32 * i));
        }
#pragma unroll
        for (int i = 0; i < NS; ++i) {
            const bool mine = w[i].x == cw.x && w[i].y == cw.y;
            if (__all_sync(0xffffffffu, mine)) {
                // fast path: the 8 entries of this word are in registers
#pragma unroll
                for (int q = 0; q < 8; ++q)
                    if (q < cnz) sum[i] = fma(cv[q], *reinterpret_cast<const double*>(xr[i] + cob[q]), sum[i]);
            } else {
                // generic path for this slice
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const unsigned int c = code_at(w[i], q);
                    if (c != 255u) sum[i] = fma(P.e[c].v, *reinterpret_cast<const double*>(xr[i] + P.e[c].offb), sum[i]);
                }
                // all lanes agree on another word: make it the cached one
                const unsigned int lx = __shfl_sync(0xffffffffu, w[i].x, 0), ly = __shfl_sync(0xffffffffu, w[i].y, 0);
                if (__all_sync(0xffffffffu, w[i].x == lx && w[i].y == ly)) {
                    cw = make_uint2(lx, ly);
                    cnz = 0;
                    bool trailing_pad_only = true;
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        const unsigned int c = code_at(cw, q);
                        cv[q] = P.e[c].v;          // entry 255 of the parameter block is zero-filled, never used (q >= cnz)
                        cob[q] = P.e[c].offb;
                        if (c != 255u) {
                            if (cnz == q) cnz = q + 1;
                            else trailing_pad_only = false;       // a code after a padding byte: the layout never produces it
                        }
                    }
                    if (!trailing_pad_only) { cw = make_uint2(~0u, ~0u); cnz = 0; }
                }
            }
        }
        const int maxb = max(nb[0], nb[1]);
        for (int b = 1; b < maxb; ++b) {          // rows with more than 8 non-zeros: generic path
            load_codes(p0, nb, b, w);
#pragma unroll
            for (int i = 0; i < NS; ++i)
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const unsigned int c = code_at(w[i], q);
                    if (c != 255u) sum[i] = fma(P.e[c].v, *reinterpret_cast<const double*>(xr[i] + P.e[c].offb), sum[i]);
                }
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
#pragma unroll
        for (int i = 0; i < NS; ++i) {
            w[i] = wn[i];
            p0[i] = p0n[i]; nb[i] = nbn[i];
            p0n[i] = p0nn[i]; nbn[i] = nbnn[i];
        }
    }
}

}  // namespace

namespace calz {

// called from launch_selld (mpk.cu) when the option is set and the constant-bank persistent kernel would have been used
int launch_selld_ufast(calz_mat* m, const double* x, const double* xp, double* y, int64_t s0, int64_t s1, bool newton, double shift,
                       double pair) {
    calz_ctx* ctx = m->ctx;
    static_assert(sizeof(DictParam) == sizeof(m->h_dict), "dictionary parameter block");
    if (s1 <= s0) return CALZ_OK;
    static int occ[2] = {0, 0};
    if (!occ[newton]) {
        if (newton) CALZ_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ[1], k_spmv_selld_ufast<true>, kSpmvThreads, 0));
        else CALZ_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ[0], k_spmv_selld_ufast<false>, kSpmvThreads, 0));
        if (occ[newton] < 1) occ[newton] = 1;
    }
    const int64_t per_cta = (int64_t)(kSpmvThreads / 32) * kDictSlicesPerWarp;
    const unsigned grid = (unsigned)std::min<int64_t>((s1 - s0 + per_cta - 1) / per_cta, (int64_t)ctx->num_sms * occ[newton]);
    const DictParam& Pd = *(const DictParam*)m->h_dict;
    if (newton)
        k_spmv_selld_ufast<true><<<grid, kSpmvThreads, 0, ctx->stream>>>(m->d_slice_ptr, (const uint2*)m->d_codes, Pd, x, xp, y, (int)s0, (int)s1,
                                                                        (int)m->n_loc, shift, pair);
    else
        k_spmv_selld_ufast<false><<<grid, kSpmvThreads, 0, ctx->stream>>>(m->d_slice_ptr, (const uint2*)m->d_codes, Pd, x, xp, y, (int)s0, (int)s1,
                                                                         (int)m->n_loc, 0.0, 0.0);
    CALZ_LAUNCH_CHECK(ctx);
    return CALZ_OK;
}

}  // namespace calz

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

// ---- persistent MPK: all s steps in ONE launch, ordered by data-flow flags so that the slice of A in use stays in L2
//
// The s SpMV steps of a block read A s times; at C3 that is 11.2 of the 13.4 GB a plain MPK moves.  A[rows r..r+s*b]
// (b = bandwidth) fits the 126 MB L2 when the steps are skewed: tile t of step k runs as soon as tiles t-G..t+G of step
// k-1 are done (G = ceil(b/tile)), in the wave-front order  key = t + (k-1)*lag, lag = G + grid/s + 2 (all s steps are
// in flight at once, each on its own range of tiles, step k trailing step k-1 by `lag` tiles).  CTAs are persistent and take the
// schedule entries round-robin (entry e -> CTA e mod grid): every dependency of an entry sits at a smaller entry, every
// CTA walks its entries in increasing order, so the smallest unfinished entry can always run -- no dead-lock, no global
// barrier, no atomics on the critical path.  Completion flags: one word per (step, tile), release/acquire at GPU scope.
constexpr int kPersistThreads = 512;
constexpr int kPersistTileSlices = 32;               // 1024 rows per tile
constexpr unsigned long long kPersistSpinLimit = 50ull * 1000ull * 1000ull;

struct PersistArgs {
    const int32_t* slice_ptr; const int32_t* col; const double* val;
    double* W; long long ldW; long long n_loc; long long nslices;
    int s, G, ntiles, nsched;
    const unsigned int* sched;                        // (k << 27) | tile, wave-front order
    unsigned int* done;                               // [s+1][ntiles]
    int* err;
    double shift[32], pair[32];
    int tlo[33], thi[33];                             // active tile range of step k
};

__device__ __forceinline__ uint32_t mpk_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mpk_mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mpk_smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mpk_mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(mpk_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mpk_mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(mpk_smem_u32(bar)), "r"(parity) : "memory");
    } while (!done);
}

// four consecutive entries j..j+3 of one lane's row inside a SELL slice (predicated past the slice width)
struct Quad { double a[4]; int32_t c[4]; };
__device__ __forceinline__ Quad load_quad(const double* __restrict__ v, const int32_t* __restrict__ c, int j, int w, int32_t self) {
    Quad q;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        const bool ok = j + u < w;
        q.a[u] = ok ? __ldg(v + (j + u) * 32) : 0.0;
        q.c[u] = ok ? __ldg(c + (j + u) * 32) : self;
    }
    return q;
}

// two slices per warp with all loads of a round in flight together (the persistent kernel has fewer warps per SM than
// the one-slice-per-warp kernel, so the memory-level parallelism has to come from the instruction stream)
template <bool NEWTON>
__device__ __forceinline__ void sell_slice_pair(const PersistArgs& a, long long s0, long long s1, int lane, const double* __restrict__ x,
                                                const double* __restrict__ xprev, double* __restrict__ y, double shift, double pair) {
    const bool has1 = s1 < a.nslices;
    const int32_t p00 = __ldg(a.slice_ptr + s0), p01 = __ldg(a.slice_ptr + s0 + 1);
    const int32_t p10 = has1 ? __ldg(a.slice_ptr + s1) : 0, p11 = has1 ? __ldg(a.slice_ptr + s1 + 1) : 0;
    const double* __restrict__ v0 = a.val + (long long)p00 * 32 + lane;
    const int32_t* __restrict__ c0 = a.col + (long long)p00 * 32 + lane;
    const double* __restrict__ v1 = a.val + (long long)p10 * 32 + lane;
    const int32_t* __restrict__ c1 = a.col + (long long)p10 * 32 + lane;
    const int w0 = p01 - p00, w1 = p11 - p10;
    const long long r0 = s0 * 32 + lane, r1 = s1 * 32 + lane;
    const int32_t self0 = (int32_t)min(r0, a.n_loc - 1), self1 = (int32_t)min(has1 ? r1 : r0, a.n_loc - 1);
    double sum0 = 0.0, sum1 = 0.0;
    const int wmax = max(w0, w1);
    for (int j = 0; j < wmax; j += 4) {
        const Quad q0 = load_quad(v0, c0, j, w0, self0);
        const Quad q1 = load_quad(v1, c1, j, w1, self1);
        double x0[4], x1[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) { x0[u] = x[q0.c[u]]; x1[u] = x[q1.c[u]]; }
#pragma unroll
        for (int u = 0; u < 4; ++u) { sum0 = fma(q0.a[u], x0[u], sum0); sum1 = fma(q1.a[u], x1[u], sum1); }
    }
    if (r0 < a.n_loc) {
        if (NEWTON) sum0 = newton_epilogue(sum0, x[r0], pair != 0.0 ? xprev[r0] : 0.0, shift, pair);
        y[r0] = sum0;
    }
    if (has1 && r1 < a.n_loc) {
        if (NEWTON) sum1 = newton_epilogue(sum1, x[r1], pair != 0.0 ? xprev[r1] : 0.0, shift, pair);
        y[r1] = sum1;
    }
}

// warp 0 = control (dependency polling, completion flags), warps 1..16 = compute; two tiles in flight per CTA
template <bool NEWTON>
__global__ void __launch_bounds__(kPersistThreads + 32, 2)
k_mpk_persistent(PersistArgs a) {
    __shared__ uint64_t ready[2], finished[2];
    __shared__ int bail;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int CW = kPersistThreads / 32;             // compute warps
    if (threadIdx.x == 0) {
        bail = 0;
        mpk_mbar_init(&ready[0], 1); mpk_mbar_init(&ready[1], 1);
        mpk_mbar_init(&finished[0], CW); mpk_mbar_init(&finished[1], CW);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (warp == 0) {
        // ======================================================= control warp
        int i = 0, prev_k = 0, prev_t = 0;
        for (int e = blockIdx.x; e < a.nsched; e += gridDim.x, ++i) {
            const unsigned int ent = a.sched[e];
            const int k = (int)(ent >> 27), t = (int)(ent & ((1u << 27) - 1));
            if (k > 1) {
                // wait for tiles [t-G, t+G] of step k-1 (clipped to that step's active range)
                const int lo = max(t - a.G, a.tlo[k - 1]), hi = min(t + a.G + 1, a.thi[k - 1]);
                const volatile unsigned int* d = a.done + (size_t)(k - 1) * a.ntiles;
                unsigned long long it = 0;
                while (true) {
                    bool ok = true;
                    for (int j = lo + lane; j < hi; j += 32) ok &= (d[j] != 0u);
                    if (__all_sync(0xffffffffu, ok)) break;
                    if (++it > kPersistSpinLimit || (it % 1024 == 0 && *(volatile int*)a.err)) {
                        if (lane == 0) { *(volatile int*)a.err = 2; bail = 1; }
                        break;
                    }
                }
                __threadfence();                       // acquire (also drops stale L1 lines of this SM)
            }
            __syncwarp();
            if (lane == 0) mpk_mbar_arrive(&ready[i & 1]);
            if (i > 0) {                               // publish the previous tile while the compute warps work on this one
                mpk_mbar_wait(&finished[(i - 1) & 1], (uint32_t)(((i - 1) >> 1) & 1));
                if (lane == 0) {
                    __threadfence();                   // release: the tile's rows before its flag
                    *((volatile unsigned int*)a.done + (size_t)prev_k * a.ntiles + prev_t) = 1u;
                }
            }
            prev_k = k; prev_t = t;
            if (bail) break;
        }
        if (i > 0 && !bail) {
            mpk_mbar_wait(&finished[(i - 1) & 1], (uint32_t)(((i - 1) >> 1) & 1));
            if (lane == 0) {
                __threadfence();
                *((volatile unsigned int*)a.done + (size_t)prev_k * a.ntiles + prev_t) = 1u;
            }
        }
    } else {
        // ======================================================= compute warps
        const int cw = warp - 1;
        int i = 0;
        for (int e = blockIdx.x; e < a.nsched; e += gridDim.x, ++i) {
            const unsigned int ent = a.sched[e];
            const int k = (int)(ent >> 27), t = (int)(ent & ((1u << 27) - 1));
            mpk_mbar_wait(&ready[i & 1], (uint32_t)((i >> 1) & 1));
            if (*(volatile int*)&bail) return;
            const double* x = a.W + (long long)(k - 1) * a.ldW;
            const double* xp = (k >= 2) ? a.W + (long long)(k - 2) * a.ldW : x;
            double* y = a.W + (long long)k * a.ldW;
            const long long s0 = (long long)t * kPersistTileSlices + cw;
            if (s0 < a.nslices) sell_slice_pair<NEWTON>(a, s0, s0 + CW, lane, x, xp, y, a.shift[k - 1], a.pair[k - 1]);
            __syncwarp();
            if (lane == 0) mpk_mbar_arrive(&finished[i & 1]);
        }
    }
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

// one SpMV step on local rows [lo,hi) (already aligned to the layout granule)
int spmv_step(calz_mat* m, const double* x, const double* xp, double* y, int64_t lo, int64_t hi, bool newton,
              double shift, double pair) {
    if (hi <= lo) return CALZ_OK;
    calz_ctx* ctx = m->ctx;
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

// the ONE level-s halo exchange of an outer step: column `col` of the workspace
int halo_exchange(calz_mat* m, double* w) {
    calz_ctx* ctx = m->ctx;
    const int P = ctx->nranks;
    if (P <= 1) return CALZ_OK;
    if (m->p2p_halo) return p2p_halo_exchange(m, w);      // push into the peers' ghost zones over NVLink (p2p.cu)
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
    CALZ_TRY(halo_exchange(m, W));

    const int64_t gran = (m->layout == CALZ_LAYOUT_SELL) ? (m->d_perm ? m->sell_sigma : 32) : 1;
    auto lo_of = [&](int k) { return (m->hull_lo[s - k] / gran) * gran; };
    auto hi_of = [&](int k) { return std::min<int64_t>(m->n_loc, round_up(m->hull_hi[s - k], gran)); };
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
    if (ctx->opt_l2_chunk_bytes > 0 && bytes_per_row > 0) {
        chunk_rows = round_up(std::max<int64_t>(ctx->opt_l2_chunk_bytes / bytes_per_row, gran), gran);
        if (chunk_rows < 2 * bwid || chunk_rows >= m->n_loc) chunk_rows = 0;   // band too wide / matrix fits: plain sweeps
    }
    // ---- persistent data-flow kernel (SELL without row permutation, wave-front window fits in L2)
    if (ctx->opt_mpk_persistent && m->layout == CALZ_LAYOUT_SELL && !m->d_perm && s <= 31) {
        const long long TR = 32LL * kPersistTileSlices;
        const int ntiles = (int)((m->n_loc + TR - 1) / TR);
        const int G = (int)((m->bandwidth + TR - 1) / TR);
        int per_sm = 0;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_mpk_persistent<true>, kPersistThreads + 32, 0);
        const int grid = std::min(ctx->num_sms * std::max(1, std::min(per_sm, 2)), std::max(1, ntiles));
        // lag between consecutive steps: the dependency reach G plus the tiles of one step that are in flight at the same
        // time (grid/s), so that everything an entry waits for sits at least `grid` schedule positions behind it
        const int lag = G + (grid + s - 1) / s + 2;
        const long long window_tiles = (long long)(s - 1) * lag + grid;
        const long long window_bytes = window_tiles * TR * bytes_per_row;
        if (per_sm >= 1 && ntiles >= 4 * grid && G + 1 < ntiles && window_bytes <= ctx->opt_mpk_l2_window_bytes) {
            PersistSched& ps = m->persist[s];
            if (!ps.d_sched) {
                std::vector<std::pair<long long, unsigned int>> ent;
                for (int k = 1; k <= s; ++k) {
                    ps.tlo[k] = (int)(lo_of(k) / TR);
                    ps.thi[k] = (int)std::min<long long>(ntiles, (hi_of(k) + TR - 1) / TR);
                    for (int t = ps.tlo[k]; t < ps.thi[k]; ++t)
                        ent.push_back({(long long)t + (long long)(k - 1) * lag, ((unsigned int)k << 27) | (unsigned int)t});
                }
                ps.tlo[0] = 0; ps.thi[0] = ntiles;
                std::stable_sort(ent.begin(), ent.end(), [](const std::pair<long long, unsigned int>& x, const std::pair<long long, unsigned int>& y) {
                    return x.first < y.first; });
                std::vector<unsigned int> sched(ent.size());
                for (size_t i = 0; i < ent.size(); ++i) sched[i] = ent[i].second;
                ps.nsched = (int)sched.size();
                CALZ_CUDA(ctx, cudaMalloc(&ps.d_sched, std::max<size_t>(sched.size(), 1) * sizeof(unsigned int)));
                CALZ_CUDA(ctx, cudaMemcpyAsync(ps.d_sched, sched.data(), sched.size() * sizeof(unsigned int), cudaMemcpyHostToDevice, ctx->stream));
                CALZ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
                CALZ_CUDA(ctx, cudaMalloc(&ps.d_done, (size_t)(s + 1) * ntiles * sizeof(unsigned int)));
                if (!m->d_persist_err) {
                    CALZ_CUDA(ctx, cudaMalloc(&m->d_persist_err, 64));
                    CALZ_CUDA(ctx, cudaMemset(m->d_persist_err, 0, 64));
                }
            }
            CALZ_CUDA(ctx, cudaMemsetAsync(ps.d_done, 0, (size_t)(s + 1) * ntiles * sizeof(unsigned int), ctx->stream));
            PersistArgs a{};
            a.slice_ptr = m->d_slice_ptr; a.col = m->d_sell_col; a.val = m->d_sell_val;
            a.W = W; a.ldW = ld; a.n_loc = m->n_loc; a.nslices = m->sell_slices;
            a.s = s; a.G = G; a.ntiles = ntiles; a.nsched = ps.nsched; a.sched = ps.d_sched; a.done = ps.d_done; a.err = m->d_persist_err;
            for (int k = 0; k < s; ++k) { a.shift[k] = sh.re[k]; a.pair[k] = sh.pair[k]; }
            for (int k = 0; k <= s; ++k) { a.tlo[k] = ps.tlo[k]; a.thi[k] = ps.thi[k]; }
            if (sh.newton) k_mpk_persistent<true><<<grid, kPersistThreads + 32, 0, ctx->stream>>>(a);
            else k_mpk_persistent<false><<<grid, kPersistThreads + 32, 0, ctx->stream>>>(a);
            CALZ_LAUNCH_CHECK(ctx);
            if (m->p2p_halo) CALZ_TRY(p2p_halo_ack(m));
            return CALZ_OK;
        }
    }
    if (chunk_rows == 0) {
        for (int k = 1; k <= s; ++k) CALZ_TRY(step(k, 0, m->n_loc));
    } else {
        const int64_t span = m->n_loc + (int64_t)(s - 1) * bwid;
        for (int64_t c0 = 0; c0 < span; c0 += chunk_rows)
            for (int k = 1; k <= s; ++k) {
                const int64_t lo = c0 - (int64_t)(k - 1) * bwid, hi = lo + chunk_rows;
                if (hi <= 0 || lo >= m->n_loc) continue;
                CALZ_TRY(step(k, std::max<int64_t>(lo, 0), std::min<int64_t>(hi, m->n_loc)));
            }
    }
    if (m->p2p_halo) CALZ_TRY(p2p_halo_ack(m));           // column-0 ghosts consumed: the owners may push the next ones
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

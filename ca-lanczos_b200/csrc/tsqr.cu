// K5/K6/K7: TSQR -- per-row-block Householder QR and a reduction tree.
//
// Reference counterpart: tsqr.m:7-12 ([Q,R]=qr(A,0); d=sign(diag(R)); R=diag(d)*R; Q=Q*diag(d)).  MATLAB's qr is LAPACK
// Householder (geqrf+orgqr); this is the same factorisation organised as a tree:
//   level 0   a CTA of 256 threads owns a leaf of 256*RPL rows, RPL rows per thread held in registers (1024 rows for c <= 8),
//             and reduces it to R (c x c) with c Householder reflectors (dlarfg/dlarf arithmetic), keeps the reflectors;
//   level l   the leaf R factors, stacked as a tall (leaves*c) x c matrix, are reduced by the SAME kernel until one R is left
//             (16.7 M rows x 8: 16384 -> 128 -> 1 leaves); with a communicator the per-rank R factors are gathered and
//             reduced redundantly on every rank (deterministic => identical R everywhere);
//   top-down  Q = Q_0 Q_1 ... Q_top [D;0]: the same walk backwards, each leaf applying its reflectors to the c x c slice
//             handed down by its parent.  D = sign(diag(R)) is the reference's sign fix (:9-11), sign(0)=0 included.
//
// Why a CTA-wide leaf: a Householder column costs one norm/dot reduction round trip (~0.3 us of dependent latency) whatever
// the leaf size, so the bytes a leaf brings per round trip decide whether the sweep can keep up with HBM (round 1: one warp
// per 256-row leaf, two round trips per column, 0.20 of the copy peak).  Here ONE batched reduction per column carries the
// sub-column norm and the dot products with all trailing columns at once: a butterfly reduce-scatter inside each warp
// (values halve as the distance halves: 9 shuffles for 8 values instead of 40), one shared-memory exchange between the 8
// warps, ONE __syncthreads per column (the exchange buffers alternate).
#include <algorithm>
#include <type_traits>

#include "tsops.cuh"

namespace calz {

namespace {

constexpr int kTsqrThreads = 128;                // 4 warps per leaf; RPL rows per thread (RPL*CW <= 64 doubles of the leaf per thread)
constexpr int kTsqrWarps = kTsqrThreads / 32;
#ifndef CALZ_TSQR_LEAF_CTAS8
#define CALZ_TSQR_LEAF_CTAS8 3
#endif
constexpr int kLeafCtas8 = CALZ_TSQR_LEAF_CTAS8;      // resident leaf CTAs per SM for c <= 8 (3: <= 168 registers per thread)

// compile-time loop: f(std::integral_constant<int, I>) for I = B .. E-1
template <int B, int E, class F>
__device__ __forceinline__ void static_for(F&& f) {
    if constexpr (B < E) {
        f(std::integral_constant<int, B>{});
        static_for<B + 1, E>(f);
    }
}

constexpr int pow2_ceil(int k) { return k <= 1 ? 1 : (k <= 2 ? 2 : (k <= 4 ? 4 : (k <= 8 ? 8 : (k <= 16 ? 16 : 32)))); }

// Butterfly reduce-scatter of KP (power of two <= 32) values over the 32 lanes.  On return p[0] holds the warp total of value
// number idx = lane >> log2(32/KP) (the lanes that share an idx all hold it).
template <int KP, int OFF = 16>
__device__ __forceinline__ void warp_reduce_scatter(double (&p)[KP], int lane) {
    if constexpr (OFF >= 1) {
        if constexpr (KP >= 2) {
            constexpr int H = KP / 2;
            const bool up = (lane & OFF) != 0;
            double q[H];
#pragma unroll
            for (int k = 0; k < H; ++k) {
                const double send = up ? p[k] : p[k + H];
                const double keep = up ? p[k + H] : p[k];
                q[k] = keep + __shfl_xor_sync(0xffffffffu, send, OFF);
            }
            warp_reduce_scatter<H, OFF / 2>(q, lane);
            p[0] = q[0];
        } else {
            p[0] += __shfl_xor_sync(0xffffffffu, p[0], OFF);
            warp_reduce_scatter<1, OFF / 2>(p, lane);
        }
    }
}

// every thread contributes p[0..KP); afterwards red[w*KP + k], w = 0..3, are the per-warp totals of value k
template <int KP>
__device__ __forceinline__ void cta_reduce_post(double (&p)[KP], double* red, int lane, int warp) {
    warp_reduce_scatter<KP>(p, lane);
    constexpr int SH = 32 / KP;                       // lanes per value
    if ((lane & (SH - 1)) == 0) red[warp * KP + lane / SH] = p[0];
}
__device__ __forceinline__ double cta_reduce_get(const double* red, int KP, int k) {
    double s = 0.0;
#pragma unroll
    for (int w = 0; w < kTsqrWarps; ++w) s += red[w * KP + k];
    return s;
}

// ---------------------------------------------------------------------------------------------------------------- factor
// One leaf per CTA: local row lr = i*128 + tid (i < RPL).  Column j: pivot = local row j (thread j, i = 0).  Per column ONE batched
// reduction of CW values over this thread's rows BELOW the pivot:
//     sum x^2 | sum x*a_l for the trailing columns l > j | sum x*v_i for the finished reflectors i < j
// -> butterfly -> shared memory; barrier; warp 0 finishes the sums, runs dlarfg ONCE (sqrt and the two reciprocals are ~100
// instructions in fp64: not worth repeating in 128 threads), posts {beta, tau, scale} and per trailing column {w_l, w_l*scale}, and
// keeps v_i'v_j (the compact-WY data the top-down sweep needs, see k_tsqr_apply); barrier; every thread scales its part of the
// reflector and updates its rows.
template <int CW, int RPL>
__global__ void __launch_bounds__(kTsqrThreads, (CW <= 8 ? kLeafCtas8 : 2))
k_tsqr_leaf(long long nrows, int c, const double* A, long long ldA, double* V, long long ldV, double* __restrict__ tau_out,
            double* __restrict__ Rstack, long long ldR, double* __restrict__ Gout, const int* __restrict__ pred, int want) {
    if (pred && *pred != want) return;
    constexpr int LEAF = kTsqrThreads * RPL;
    constexpr int KP = pow2_ceil(CW);
    __shared__ double red[kTsqrWarps * KP];
    __shared__ double piv[CW];                        // the pivot row: columns j.. at [0, K), columns 0..j-1 at [K, CW)
    __shared__ __align__(16) double coef[2 * CW];     // {w_l, w_l*scale} per trailing column
    __shared__ double hdr[4];                         // beta, tau, scale
    __shared__ double Gs[CW * CW];                    // strict upper triangle of V'V
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const long long r0 = (long long)blockIdx.x * LEAF;

    double a[RPL][CW];
#pragma unroll
    for (int i = 0; i < RPL; ++i) {
        const long long row = r0 + i * kTsqrThreads + tid;
#pragma unroll
        for (int j = 0; j < CW; ++j) a[i][j] = (row < nrows && j < c) ? A[row + (long long)j * ldA] : 0.0;
    }
    for (int e = tid; e < CW * CW; e += kTsqrThreads) Gs[e] = 0.0;
    double my_tau = 0.0;                              // thread j keeps tau_j

    static_for<0, CW>([&](auto jc) {
        constexpr int j = decltype(jc)::value;
        if (j < c) {
            constexpr int K = CW - j;
            // value k < K: column j + k ; value K + i: finished reflector i < j
            double p[KP];
#pragma unroll
            for (int k = 0; k < KP; ++k) p[k] = 0.0;
            double x[RPL];
#pragma unroll
            for (int i = 0; i < RPL; ++i) {
                x[i] = (i > 0 || tid > j) ? a[i][j] : 0.0;
#pragma unroll
                for (int k = 0; k < K; ++k) p[k] = fma(x[i], a[i][j + k], p[k]);
#pragma unroll
                for (int q = 0; q < j; ++q) p[K + q] = fma(x[i], a[i][q], p[K + q]);
            }
            cta_reduce_post<KP>(p, red, lane, warp);
            if (tid == j) {                           // the pivot row travels through shared memory as well
#pragma unroll
                for (int k = 0; k < K; ++k) piv[k] = a[0][j + k];
#pragma unroll
                for (int q = 0; q < j; ++q) piv[K + q] = a[0][q];
            }
            __syncthreads();
            if (warp == 0) {
                // dlarfg on column j, then w_l = tau * v'a_l = tau * (pivot entry + scale * sum_below x*a_l)
                const double S = (lane < CW) ? cta_reduce_get(red, KP, lane) : 0.0;
                const double ss = __shfl_sync(0xffffffffu, S, 0);
                const double alpha = piv[0];
                double tj = 0.0, scale = 0.0, beta = alpha;
                if (ss != 0.0) {
                    beta = -copysign(sqrt(fma(alpha, alpha, ss)), alpha);
                    tj = (beta - alpha) / beta;
                    scale = 1.0 / (alpha - beta);
                }
                if (lane >= 1 && lane < K) {
                    const double w = tj * fma(scale, S, piv[lane]);
                    coef[2 * lane] = w;
                    coef[2 * lane + 1] = w * scale;
                }
                if (lane >= K && lane < CW) Gs[(lane - K) * CW + j] = fma(scale, S, piv[lane]);      // v_i'v_j, i = lane - K
                if (lane == 0) { hdr[0] = beta; hdr[1] = tj; hdr[2] = scale; }
            }
            __syncthreads();
            const double beta = hdr[0], tj = hdr[1], scale = hdr[2];
            if (tid == j) my_tau = tj;
            // v: 1 at the pivot (beta is stored there), x*scale below; rows above keep their R entries
#pragma unroll
            for (int i = 0; i < RPL; ++i) {
                if (i > 0 || tid > j) a[i][j] = x[i] * scale;
                else if (tid == j) a[i][j] = beta;
            }
            if (tj != 0.0) {
#pragma unroll
                for (int k = 1; k < K; ++k) {
                    if (j + k < c) {
                        const double2 wc = *reinterpret_cast<const double2*>(&coef[2 * k]);
                        // below rows: a_l -= (w*scale) * x ; pivot row: a_l -= w   (x is 0 on the pivot row and above it)
#pragma unroll
                        for (int i = 0; i < RPL; ++i) a[i][j + k] = fma(-wc.y, x[i], a[i][j + k]);
                        if (tid == j) a[0][j + k] -= wc.x;
                    }
                }
            }
        }
    });

    // reflectors (+R above them) back to V, R to the stack of the next level, tau, V'V
#pragma unroll
    for (int i = 0; i < RPL; ++i) {
        const long long row = r0 + i * kTsqrThreads + tid;
#pragma unroll
        for (int j = 0; j < CW; ++j)
            if (row < nrows && j < c) V[row + (long long)j * ldV] = a[i][j];
    }
    if (tid < c) {
#pragma unroll
        for (int j = 0; j < CW; ++j)
            if (j < c) Rstack[(long long)blockIdx.x * c + tid + (long long)j * ldR] = (j >= tid) ? a[0][j] : 0.0;
        tau_out[(long long)blockIdx.x * c + tid] = my_tau;
    }
    __syncthreads();
    for (int e = tid; e < c * c; e += kTsqrThreads) Gout[(long long)blockIdx.x * c * c + e] = Gs[(e / c) * CW + (e % c)];     // row-major c x c
}

// ---------------------------------------------------------------------------------------------------------------- apply
// Q_leaf = H_1 ... H_c [W_leaf; 0] for every leaf, in compact-WY form: H_1...H_c = I - V T V' with T^-1 = striu(V'V) + diag(1/tau),
// and since only the top c rows of [W;0] are non-zero,
//     Q_leaf = [W;0] - V * M,   M = T * (Vtop' * W)           (Vtop = the unit lower triangular top c x c block of V).
// striu(V'V) comes from the factorisation (no reduction here at all): a c x c triangular solve per leaf in the prologue, then
// a streaming update, RC rows per thread at a time.  A reflector with tau = 0 is the identity: its column of V is ignored.
template <int CW, int RPL, int RC>
__global__ void __launch_bounds__(kTsqrThreads, (CW <= 16 ? 4 : 3))
k_tsqr_apply(long long nrows, int c, const double* V, long long ldV, const double* __restrict__ tau_in, const double* __restrict__ G,
             const double* __restrict__ W, long long ldW, double* Out, long long ldOut) {
    static_assert(RPL % RC == 0, "row chunks");
    constexpr int LEAF = kTsqrThreads * RPL;
    __shared__ double U[CW][CW + 1];                  // striu(V'V)
    __shared__ double Vt[CW][CW + 1];                 // Vtop
    __shared__ double Ws[CW][CW + 1];
    __shared__ __align__(16) double Ms[CW][CW];
    __shared__ double taus[CW];
    const int tid = threadIdx.x;
    const long long r0 = (long long)blockIdx.x * LEAF;

    auto load_chunk = [&](int ch, double (&v)[RC][CW]) {
#pragma unroll
        for (int i = 0; i < RC; ++i) {
            const long long row = r0 + (long long)(ch * RC + i) * kTsqrThreads + tid;
#pragma unroll
            for (int j = 0; j < CW; ++j) v[i][j] = (row < nrows && j < c) ? V[row + (long long)j * ldV] : 0.0;
        }
    };
    double v[RC][CW];
    load_chunk(0, v);                                 // in flight behind the prologue
    if (tid < CW) taus[tid] = tid < c ? tau_in[(long long)blockIdx.x * c + tid] : 0.0;
    for (int e = tid; e < CW * CW; e += kTsqrThreads) {
        const int r = e % CW, l = e / CW;
        const bool in = r < c && l < c;
        Ws[r][l] = in ? W[(long long)blockIdx.x * c + r + (long long)l * ldW] : 0.0;
        U[l][r] = in ? G[(long long)blockIdx.x * c * c + (long long)l * c + r] : 0.0;          // U[i][j] = G(i,j), row-major in G
        // Vtop(r, l): 1 on the diagonal, the stored tail below it
        const long long row = r0 + r;
        Vt[r][l] = (r > l && in && row < nrows) ? V[row + (long long)l * ldV] : (r == l ? 1.0 : 0.0);
    }
    __syncthreads();
    if (tid < CW) {
        // column l = tid of  Y = Vtop' * W  and of  M = T * Y  (back substitution with T^-1 = U + diag(1/tau))
        const int l = tid;
        double m[CW];
#pragma unroll
        for (int i = 0; i < CW; ++i) {
            double y = 0.0;
#pragma unroll
            for (int r = 0; r < CW; ++r)
                if (r >= i) y = fma(Vt[r][i], Ws[r][l], y);
            m[i] = y;
        }
#pragma unroll
        for (int ii = 0; ii < CW; ++ii) {
            const int i = CW - 1 - ii;
            double acc = m[i];
#pragma unroll
            for (int k = 0; k < CW; ++k)
                if (k > i) acc = fma(-U[i][k], m[k], acc);
            m[i] = acc * taus[i];                     // tau = 0: the row is zero (identity reflector)
        }
#pragma unroll
        for (int i = 0; i < CW; ++i) Ms[i][l] = m[i];
    }
    __syncthreads();
    // Q rows = [W;0] - V*M
#pragma unroll
    for (int ch = 0; ch < RPL / RC; ++ch) {
        if (ch > 0) load_chunk(ch, v);
        if (ch == 0 && tid < c) {                     // the top c rows: unit diagonal, nothing above it
#pragma unroll
            for (int j = 0; j < CW; ++j) v[0][j] = Vt[tid][j];
        }
#pragma unroll
        for (int l0 = 0; l0 < CW; l0 += 4) {
            if (l0 < c) {
                double b[RC][4];
#pragma unroll
                for (int i = 0; i < RC; ++i)
#pragma unroll
                    for (int l = 0; l < 4; ++l) b[i][l] = (ch == 0 && i == 0 && tid < c) ? Ws[tid < CW ? tid : 0][l0 + l] : 0.0;
#pragma unroll
                for (int j = 0; j < CW; ++j) {
                    if (j < c) {
                        const double2 m0 = *reinterpret_cast<const double2*>(&Ms[j][l0]);
                        const double2 m1 = *reinterpret_cast<const double2*>(&Ms[j][l0 + 2]);
#pragma unroll
                        for (int i = 0; i < RC; ++i) {
                            b[i][0] = fma(-v[i][j], m0.x, b[i][0]);
                            b[i][1] = fma(-v[i][j], m0.y, b[i][1]);
                            b[i][2] = fma(-v[i][j], m1.x, b[i][2]);
                            b[i][3] = fma(-v[i][j], m1.y, b[i][3]);
                        }
                    }
                }
#pragma unroll
                for (int i = 0; i < RC; ++i) {
                    const long long row = r0 + (long long)(ch * RC + i) * kTsqrThreads + tid;
#pragma unroll
                    for (int l = 0; l < 4; ++l)
                        if (row < nrows && l0 + l < c) Out[row + (long long)(l0 + l) * ldOut] = b[i][l];
                }
            }
        }
    }
}

// top of the tree: R_out = diag(d) R, D = diag(d) with d = sign(diag(R))   (tsqr.m:9-11)
__global__ void k_tsqr_finish(int c, const double* __restrict__ Rtop, long long ldR, double* __restrict__ R_out,
                              double* __restrict__ D, const int* __restrict__ pred, int want) {
    if (pred && *pred != want) return;
    for (int e = threadIdx.x; e < c * c; e += blockDim.x) {
        const int i = e % c, j = e / c;
        const double rii = Rtop[i + (long long)i * ldR];
        const double d = (rii > 0.0) ? 1.0 : (rii < 0.0 ? -1.0 : 0.0);
        R_out[e] = (j >= i) ? d * Rtop[i + (long long)j * ldR] : 0.0;
        D[e] = (i == j) ? d : 0.0;
    }
}

// place this rank's c x c R into its slot of a zeroed (P*c) x c staging stack (then all-reduced = all-gathered)
__global__ void k_tsqr_slot(int c, int P, int rank, const double* __restrict__ Rloc, long long ldR, double* __restrict__ S,
                            const int* __restrict__ pred, int want) {
    const bool run = !(pred && *pred != want);
    for (int e = threadIdx.x; e < P * c * c; e += blockDim.x) {
        const int row = e % (P * c), j = e / (P * c);
        const int q = row / c, i = row % c;
        // a predicated-off call still has to feed identical data (zeros) into the collective
        S[e] = (run && q == rank) ? Rloc[i + (long long)j * ldR] : 0.0;
    }
}

struct Level {
    double* V; long long ldV; long long nrows; long long leaves; double* tau; double* Rstack; long long ldR; double* G;
    const double* src; long long ldsrc;      // what the leaf kernel reads (level 0: the caller's matrix; first global level: the staging stack)
};

struct Plan {
    std::vector<Level> levels;
    int c = 0, rpl = 0;
    long long n = 0;
    double* D = nullptr;       // c x c sign matrix handed to the top leaf
    int first_global = -1;     // index of the first level that works on the gathered stack
};

int rpl_for(int c) { return c <= 8 ? 8 : (c <= 16 ? 4 : (c <= 20 ? 3 : 2)); }      // RPL * CW <= 64 doubles of the leaf per thread

template <int CW, int RPL>
int run_leaf(calz_ctx* ctx, const Level& L, int c, const int* pred, int want) {
    k_tsqr_leaf<CW, RPL><<<(unsigned)L.leaves, kTsqrThreads, 0, ctx->stream>>>(L.nrows, c, L.src, L.ldsrc, L.V, L.ldV, L.tau, L.Rstack,
                                                                               L.ldR, L.G, pred, want);
    CALZ_LAUNCH_CHECK(ctx);
    return CALZ_OK;
}

template <int CW, int RPL>
int run_apply(calz_ctx* ctx, const Level& L, int c, const double* W, long long ldW, double* Out, long long ldOut) {
    constexpr int RC = CW <= 8 ? 4 : (CW <= 16 ? 2 : 1);
    k_tsqr_apply<CW, RPL, RC><<<(unsigned)L.leaves, kTsqrThreads, 0, ctx->stream>>>(L.nrows, c, L.V, L.ldV, L.tau, L.G, W, ldW, Out, ldOut);
    CALZ_LAUNCH_CHECK(ctx);
    return CALZ_OK;
}

int dispatch_leaf(calz_ctx* ctx, const Level& L, int c, const int* pred, int want) {
    if (c <= 8) return run_leaf<8, 8>(ctx, L, c, pred, want);
    if (c <= 12) return run_leaf<12, 4>(ctx, L, c, pred, want);
    if (c <= 16) return run_leaf<16, 4>(ctx, L, c, pred, want);
    if (c <= 20) return run_leaf<20, 3>(ctx, L, c, pred, want);
    if (c <= 24) return run_leaf<24, 2>(ctx, L, c, pred, want);
    return run_leaf<32, 2>(ctx, L, c, pred, want);
}

int dispatch_apply(calz_ctx* ctx, const Level& L, int c, const double* W, long long ldW, double* Out, long long ldOut) {
    if (c <= 8) return run_apply<8, 8>(ctx, L, c, W, ldW, Out, ldOut);
    if (c <= 12) return run_apply<12, 4>(ctx, L, c, W, ldW, Out, ldOut);
    if (c <= 16) return run_apply<16, 4>(ctx, L, c, W, ldW, Out, ldOut);
    if (c <= 20) return run_apply<20, 3>(ctx, L, c, W, ldW, Out, ldOut);
    if (c <= 24) return run_apply<24, 2>(ctx, L, c, W, ldW, Out, ldOut);
    return run_apply<32, 2>(ctx, L, c, W, ldW, Out, ldOut);
}

Plan* plan_of(calz_ctx* ctx) {
    if (!ctx->tsqr_plan) ctx->tsqr_plan = new Plan();
    return (Plan*)ctx->tsqr_plan;
}

}  // namespace

void tsqr_plan_free(calz_ctx* ctx) {
    delete (Plan*)ctx->tsqr_plan;
    ctx->tsqr_plan = nullptr;
}

int tsqr_factor(calz_ctx* ctx, int64_t n, int c, const double* A, int64_t ldA, double* R_dev, const int* pred, int want) {
    if (c < 1 || c > kMaxC) return set_error(ctx, CALZ_ERR_UNSUPPORTED, "tsqr: c=%d outside [1,%d]", c, kMaxC);
    const int P = ctx->nranks;
    const int rpl = rpl_for(c);
    const long long LEAF = (long long)kTsqrThreads * rpl;
    // ---- plan the levels and carve the scratch
    std::vector<long long> rows, leaves;
    long long nr = n;
    int first_global = -1;
    while (true) {
        long long lv = std::max<long long>((nr + LEAF - 1) / LEAF, 1);
        rows.push_back(nr);
        leaves.push_back(lv);
        if (lv == 1) {
            if (P > 1 && first_global < 0) {       // local tree done: continue on the gathered stack
                first_global = (int)rows.size();
                nr = (long long)P * c;
                continue;
            }
            break;
        }
        nr = lv * c;
    }
    const int nlev = (int)rows.size();
    size_t doubles = 2 * (size_t)c * c;                         // D, spare
    for (int l = 0; l < nlev; ++l) doubles += 2 * (size_t)leaves[l] * c * c + (size_t)leaves[l] * c + 8;
    if (P > 1) doubles += 2 * (size_t)P * c * c;                // staging stack (collective buffer) + the global level's reflectors
    const long long ldV0 = round_up(n, 32);
    CALZ_TRY(reserve(ctx, ctx->work[0], (size_t)ldV0 * c * sizeof(double)));
    CALZ_TRY(reserve(ctx, ctx->tsqr_r, doubles * sizeof(double)));
    double* p = (double*)ctx->tsqr_r.p;
    Plan& plan = *plan_of(ctx);
    plan.levels.assign(nlev, Level{});
    plan.c = c; plan.rpl = rpl; plan.n = n; plan.first_global = first_global;
    plan.D = p; p += 2 * (size_t)c * c;
    // The staging stack is what the collective writes; the reflectors of the global level live in their OWN buffer: a
    // predicated-off factorisation (second pass of projectAndNormalize not firing) still runs the collective -- on zeros --
    // and must not disturb the reflectors of the factorisation that tsqr_form_q is going to apply.
    double *stage = nullptr, *gV = nullptr;
    if (P > 1) { stage = p; p += (size_t)P * c * c; gV = p; p += (size_t)P * c * c; }
    for (int l = 0; l < nlev; ++l) {
        Level& L = plan.levels[l];
        L.nrows = rows[l];
        L.leaves = leaves[l];
        L.Rstack = p; p += (size_t)leaves[l] * c * c;
        L.ldR = leaves[l] * c;
        L.tau = p; p += (size_t)leaves[l] * c + 8;
        L.G = p; p += (size_t)leaves[l] * c * c;
    }
    for (int l = 0; l < nlev; ++l) {
        Level& L = plan.levels[l];
        if (l == 0) { L.V = (double*)ctx->work[0].p; L.ldV = ldV0; L.src = A; L.ldsrc = ldA; }
        else if (l == first_global) { L.V = gV; L.ldV = (long long)P * c; L.src = stage; L.ldsrc = (long long)P * c; }
        else { L.V = plan.levels[l - 1].Rstack; L.ldV = plan.levels[l - 1].ldR; L.src = L.V; L.ldsrc = L.ldV; }     // reflectors overwrite the stack
    }
    // ---- bottom-up
    for (int l = 0; l < nlev; ++l) {
        const Level& L = plan.levels[l];
        if (l == first_global) {
            const Level& T = plan.levels[l - 1];
            k_tsqr_slot<<<1, 256, 0, ctx->stream>>>(c, P, ctx->rank, T.Rstack, T.ldR, stage, pred, want);
            CALZ_LAUNCH_CHECK(ctx);
            CALZ_TRY(allreduce_sum(ctx, stage, (size_t)P * c * c));
        }
        CALZ_TRY(dispatch_leaf(ctx, L, c, pred, want));
    }
    const Level& T = plan.levels[nlev - 1];
    k_tsqr_finish<<<1, 256, 0, ctx->stream>>>(c, T.Rstack, T.ldR, R_dev, plan.D, pred, want);
    CALZ_LAUNCH_CHECK(ctx);
    return CALZ_OK;
}

int tsqr_form_q(calz_ctx* ctx, int64_t n, int c, const double* A, int64_t ldA, double* Q, int64_t ldQ) {
    (void)A; (void)ldA;    // the reflectors of the last tsqr_factor that ran are on the device
    Plan& plan = *plan_of(ctx);
    if (plan.levels.empty() || plan.n != n || plan.c != c)
        return set_error(ctx, CALZ_ERR_BADARG, "tsqr_form_q: no matching factorisation");
    const int nlev = (int)plan.levels.size();
    const double* W = plan.D;
    long long ldW = c;
    for (int l = nlev - 1; l >= 0; --l) {
        const Level& L = plan.levels[l];
        double* Out = (l == 0) ? Q : L.V;
        const long long ldOut = (l == 0) ? ldQ : L.ldV;
        CALZ_TRY(dispatch_apply(ctx, L, c, W, ldW, Out, ldOut));
        W = Out;
        ldW = ldOut;
        if (l == plan.first_global) {     // hand the local top leaf this rank's c x c slice of the global walk
            W = Out + (long long)ctx->rank * c;
        }
    }
    return CALZ_OK;
}

}  // namespace calz

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

constexpr int kTsqrThreads = 256;
constexpr int kTsqrWarps = kTsqrThreads / 32;

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

// CTA-wide sums of K <= KP values: every thread contributes p[0..KP) and afterwards reads total k from red[w][k], w = 0..7
// (summed in warp order by the caller).  `red` is this column's exchange buffer: [kTsqrWarps][KP] doubles.
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
// One leaf per CTA: local row lr = i*256 + tid (i < RPL).  Column j: pivot = local row j (thread j, i = 0).
template <int CW, int RPL>
__global__ void __launch_bounds__(kTsqrThreads, 2)
k_tsqr_leaf(long long nrows, int c, const double* A, long long ldA, double* V, long long ldV, double* __restrict__ tau_out,
            double* __restrict__ Rstack, long long ldR, const int* __restrict__ pred, int want) {
    if (pred && *pred != want) return;
    constexpr int LEAF = kTsqrThreads * RPL;
    constexpr int KPMAX = pow2_ceil(CW);
    __shared__ double red[2][kTsqrWarps * KPMAX];
    __shared__ double piv[2][CW];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const long long r0 = (long long)blockIdx.x * LEAF;

    double a[RPL][CW];
#pragma unroll
    for (int i = 0; i < RPL; ++i) {
        const long long row = r0 + i * kTsqrThreads + tid;
#pragma unroll
        for (int j = 0; j < CW; ++j) a[i][j] = (row < nrows && j < c) ? A[row + (long long)j * ldA] : 0.0;
    }
    double my_tau = 0.0;                              // thread j keeps tau_j

    static_for<0, CW>([&](auto jc) {
        constexpr int j = decltype(jc)::value;
        if (j < c) {
            constexpr int K = CW - j, KP = pow2_ceil(K);
            double* rb = red[j & 1];
            double* pb = piv[j & 1];
            // partial sums over this thread's rows BELOW the pivot: p[0] -> sum x^2, p[k] -> sum x*a(:,j+k)
            double p[KP];
#pragma unroll
            for (int k = 0; k < KP; ++k) p[k] = 0.0;
#pragma unroll
            for (int i = 0; i < RPL; ++i) {
                const bool below = (i > 0) || (tid > j);
                const double x = below ? a[i][j] : 0.0;
#pragma unroll
                for (int k = 0; k < K; ++k) p[k] = fma(x, a[i][j + k], p[k]);
            }
            cta_reduce_post<KP>(p, rb, lane, warp);
            if (tid == j) {                           // the pivot row travels through shared memory as well
#pragma unroll
                for (int k = 0; k < K; ++k) pb[k] = a[0][j + k];
            }
            __syncthreads();
            // dlarfg on column j (every thread, redundantly, from identical inputs)
            const double ss = cta_reduce_get(rb, KP, 0);
            const double alpha = pb[0];
            double tj = 0.0, scale = 0.0, beta = alpha;
            if (ss != 0.0) {
                beta = -copysign(sqrt(fma(alpha, alpha, ss)), alpha);
                tj = (beta - alpha) / beta;
                scale = 1.0 / (alpha - beta);
            }
            if (tid == j) my_tau = tj;
            // v: 1 at the pivot, x*scale below, 0 above
            double v[RPL];
#pragma unroll
            for (int i = 0; i < RPL; ++i) {
                const bool below = (i > 0) || (tid > j);
                const bool pivot = (i == 0) && (tid == j);
                v[i] = below ? a[i][j] * scale : (pivot ? 1.0 : 0.0);
                if (below) a[i][j] = v[i];
                if (pivot) a[i][j] = beta;
            }
            if (tj != 0.0) {
#pragma unroll
                for (int k = 1; k < K; ++k) {
                    if (j + k < c) {
                        // w = tau * v'a_l = tau * (pivot entry + scale * sum_below x*a_l)
                        const double w = tj * fma(scale, cta_reduce_get(rb, KP, k), pb[k]);
#pragma unroll
                        for (int i = 0; i < RPL; ++i) a[i][j + k] = fma(-w, v[i], a[i][j + k]);
                    }
                }
            }
        }
    });

    // reflectors (+R above them) back to V, R to the stack of the next level, tau
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
}

// ---------------------------------------------------------------------------------------------------------------- apply
// B = H_1 ... H_c [W_leaf; 0]  for every leaf (W_leaf = rows leaf*c.. of W), written to Out (may alias V).
template <int CW, int RPL>
__global__ void __launch_bounds__(kTsqrThreads, 2)
k_tsqr_apply(long long nrows, int c, const double* V, long long ldV, const double* __restrict__ tau_in,
             const double* __restrict__ W, long long ldW, double* Out, long long ldOut) {
    constexpr int LEAF = kTsqrThreads * RPL;
    constexpr int KP = pow2_ceil(CW);
    __shared__ double red[2][kTsqrWarps * KP];
    __shared__ double taus[CW];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const long long r0 = (long long)blockIdx.x * LEAF;
    if (tid < CW) taus[tid] = tid < c ? tau_in[(long long)blockIdx.x * c + tid] : 0.0;

    double b[RPL][CW];
#pragma unroll
    for (int i = 0; i < RPL; ++i)
#pragma unroll
        for (int j = 0; j < CW; ++j)
            b[i][j] = (i == 0 && tid < c && j < c) ? W[(long long)blockIdx.x * c + tid + (long long)j * ldW] : 0.0;

    auto load_v = [&](int j, double (&v)[RPL]) {     // reflector j: 1 at local row j, stored tail below, 0 above
#pragma unroll
        for (int i = 0; i < RPL; ++i) {
            const long long row = r0 + i * kTsqrThreads + tid;
            const int lrow = i * kTsqrThreads + tid;
            const double x = (row < nrows && j >= 0 && j < c) ? V[row + (long long)j * ldV] : 0.0;
            v[i] = (lrow > j) ? x : (lrow == j ? 1.0 : 0.0);
        }
    };
    double vn[RPL];
    load_v(c - 1, vn);
    __syncthreads();
    int buf = 0;                                      // exchange buffers alternate per EXECUTED reduction (one barrier each)
    for (int j = c - 1; j >= 0; --j) {
        double v[RPL];
#pragma unroll
        for (int i = 0; i < RPL; ++i) v[i] = vn[i];
        load_v(j - 1, vn);                            // prefetch the next reflector behind this one's reduction
        const double tj = taus[j];
        if (tj != 0.0) {                              // uniform over the CTA
            double* rb = red[buf];
            buf ^= 1;
            double p[KP];
#pragma unroll
            for (int k = 0; k < KP; ++k) p[k] = 0.0;
#pragma unroll
            for (int i = 0; i < RPL; ++i)
#pragma unroll
                for (int k = 0; k < CW; ++k) p[k] = fma(v[i], b[i][k], p[k]);
            cta_reduce_post<KP>(p, rb, lane, warp);
            __syncthreads();
#pragma unroll
            for (int k = 0; k < CW; ++k) {
                if (k < c) {
                    const double w = tj * cta_reduce_get(rb, KP, k);
#pragma unroll
                    for (int i = 0; i < RPL; ++i) b[i][k] = fma(-w, v[i], b[i][k]);
                }
            }
        }
    }
#pragma unroll
    for (int i = 0; i < RPL; ++i) {
        const long long row = r0 + i * kTsqrThreads + tid;
#pragma unroll
        for (int j = 0; j < CW; ++j)
            if (row < nrows && j < c) Out[row + (long long)j * ldOut] = b[i][j];
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
    double* V; long long ldV; long long nrows; long long leaves; double* tau; double* Rstack; long long ldR;
    const double* src; long long ldsrc;      // what the leaf kernel reads (level 0: the caller's matrix; first global level: the staging stack)
};

struct Plan {
    std::vector<Level> levels;
    int c = 0, rpl = 0;
    long long n = 0;
    double* D = nullptr;       // c x c sign matrix handed to the top leaf
    int first_global = -1;     // index of the first level that works on the gathered stack
};

int cw_for(int c) { return c <= 8 ? 8 : (c <= 12 ? 12 : (c <= 16 ? 16 : (c <= 24 ? 24 : 32))); }
int rpl_for(int c) { return c <= 8 ? 4 : (c <= 16 ? 2 : 1); }          // <= 32 doubles of the leaf per thread

template <int CW, int RPL>
int run_leaf(calz_ctx* ctx, const Level& L, int c, const int* pred, int want) {
    k_tsqr_leaf<CW, RPL><<<(unsigned)L.leaves, kTsqrThreads, 0, ctx->stream>>>(L.nrows, c, L.src, L.ldsrc, L.V, L.ldV, L.tau, L.Rstack,
                                                                               L.ldR, pred, want);
    CALZ_LAUNCH_CHECK(ctx);
    return CALZ_OK;
}

template <int CW, int RPL>
int run_apply(calz_ctx* ctx, const Level& L, int c, const double* W, long long ldW, double* Out, long long ldOut) {
    k_tsqr_apply<CW, RPL><<<(unsigned)L.leaves, kTsqrThreads, 0, ctx->stream>>>(L.nrows, c, L.V, L.ldV, L.tau, W, ldW, Out, ldOut);
    CALZ_LAUNCH_CHECK(ctx);
    return CALZ_OK;
}

int dispatch_leaf(calz_ctx* ctx, const Level& L, int c, const int* pred, int want) {
    if (c <= 8) return run_leaf<8, 4>(ctx, L, c, pred, want);
    if (c <= 12) return run_leaf<12, 2>(ctx, L, c, pred, want);
    if (c <= 16) return run_leaf<16, 2>(ctx, L, c, pred, want);
    if (c <= 24) return run_leaf<24, 1>(ctx, L, c, pred, want);
    return run_leaf<32, 1>(ctx, L, c, pred, want);
}

int dispatch_apply(calz_ctx* ctx, const Level& L, int c, const double* W, long long ldW, double* Out, long long ldOut) {
    if (c <= 8) return run_apply<8, 4>(ctx, L, c, W, ldW, Out, ldOut);
    if (c <= 12) return run_apply<12, 2>(ctx, L, c, W, ldW, Out, ldOut);
    if (c <= 16) return run_apply<16, 2>(ctx, L, c, W, ldW, Out, ldOut);
    if (c <= 24) return run_apply<24, 1>(ctx, L, c, W, ldW, Out, ldOut);
    return run_apply<32, 1>(ctx, L, c, W, ldW, Out, ldOut);
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
    for (int l = 0; l < nlev; ++l) doubles += (size_t)leaves[l] * c * c + (size_t)leaves[l] * c + 8;
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

// K5/K6/K7: TSQR -- per-row-block Householder QR with warp-shuffle reductions and a reduction tree.
//
// Reference counterpart: tsqr.m:7-12 ([Q,R]=qr(A,0); d=sign(diag(R)); R=diag(d)*R; Q=Q*diag(d)).  MATLAB's
// qr is LAPACK Householder (geqrf+orgqr); this is the same factorisation organised as a tree:
//   level 0   every warp owns a leaf of 32*RPL rows (RPL rows per lane, held in registers), reduces it to
//             R (c x c) with c Householder reflectors (dlarfg/dlarf arithmetic), keeps the reflectors;
//   level l   the leaf R factors, stacked as a tall (leaves*c) x c matrix, are reduced by the SAME kernel
//             (fan-in 32*RPL/c per level) until one R remains; with a communicator the per-rank R factors are
//             gathered and reduced redundantly on every rank (deterministic => identical R everywhere);
//   top-down  Q = Q_0 Q_1 ... Q_top [D;0]: the same walk backwards, each leaf applying its reflectors to the
//             c x c slice handed down by its parent.  D = sign(diag(R)) is the reference's sign fix (:9-11),
//             sign(0)=0 included.
#include <algorithm>
#include <map>

#include "tsops.cuh"

namespace calz {

namespace {

constexpr int kTsqrThreads = 64;      // 2 warps per CTA, one leaf per warp at a time: at ~180 registers per thread the
                                      // register file holds 5 such CTAs (10 warps) per SM, but only 2 CTAs of 128 (8 warps)

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Householder QR of one leaf held in registers: lane l owns local rows i*32+l (i < RPL).
// On exit a[][] holds R in local rows 0..c-1 (upper triangle) and the reflector tails below the diagonal.
template <int CW, int RPL>
__device__ __forceinline__ void leaf_qr(double (&a)[RPL][CW], double (&tau)[CW], int c, int lane) {
#pragma unroll
    for (int j = 0; j < CW; ++j) {
        tau[j] = 0.0;
        if (j < c) {
            // dlarfg on column j, pivot = local row j (slot 0, lane j)
            double ss = 0.0;
#pragma unroll
            for (int i = 0; i < RPL; ++i) {
                const bool below = (i > 0) || (lane > j);
                ss += below ? a[i][j] * a[i][j] : 0.0;
            }
            ss = warp_sum(ss);
            const double alpha = __shfl_sync(0xffffffffu, a[0][j], j);
            double tj = 0.0, scale = 0.0, beta = alpha;
            if (ss != 0.0) {
                beta = -copysign(sqrt(alpha * alpha + ss), alpha);
                tj = (beta - alpha) / beta;
                scale = 1.0 / (alpha - beta);
            }
            tau[j] = tj;
            // v: 1 at the pivot, x*scale below, 0 above
            double v[RPL];
#pragma unroll
            for (int i = 0; i < RPL; ++i) {
                const bool below = (i > 0) || (lane > j);
                const bool pivot = (i == 0) && (lane == j);
                v[i] = below ? a[i][j] * scale : (pivot ? 1.0 : 0.0);
                if (below) a[i][j] = v[i];
                if (pivot) a[i][j] = beta;
            }
            if (tj != 0.0) {
#pragma unroll
                for (int l = j + 1; l < CW; ++l) {
                    if (l < c) {
                        double w = 0.0;
#pragma unroll
                        for (int i = 0; i < RPL; ++i) w = fma(v[i], a[i][l], w);
                        w = warp_sum(w) * tj;
#pragma unroll
                        for (int i = 0; i < RPL; ++i) a[i][l] = fma(-w, v[i], a[i][l]);
                    }
                }
            }
        }
    }
}

template <int CW, int RPL>
__global__ void __launch_bounds__(kTsqrThreads)
k_tsqr_leaf(long long nrows, int c, const double* A, long long ldA, double* V, long long ldV, double* __restrict__ tau_out,
            double* __restrict__ Rstack, long long ldR, const int* __restrict__ pred, int want) {
    if (pred && *pred != want) return;
    constexpr int LEAF = 32 * RPL;
    const int lane = threadIdx.x & 31;
    const long long nleaves = (nrows + LEAF - 1) / LEAF;
    const long long wstride = (long long)gridDim.x * (kTsqrThreads / 32);
    for (long long leaf = (long long)blockIdx.x * (kTsqrThreads / 32) + (threadIdx.x >> 5); leaf < nleaves; leaf += wstride) {
        const long long r0 = leaf * LEAF;
        double a[RPL][CW], tau[CW];
#pragma unroll
        for (int i = 0; i < RPL; ++i) {
            const long long row = r0 + i * 32 + lane;
#pragma unroll
            for (int j = 0; j < CW; ++j) a[i][j] = (row < nrows && j < c) ? A[row + (long long)j * ldA] : 0.0;
        }
        leaf_qr<CW, RPL>(a, tau, c, lane);
        // reflectors (+R above them) back to V, R to the stack of the next level, tau
#pragma unroll
        for (int i = 0; i < RPL; ++i) {
            const long long row = r0 + i * 32 + lane;
#pragma unroll
            for (int j = 0; j < CW; ++j)
                if (row < nrows && j < c) V[row + (long long)j * ldV] = a[i][j];
        }
        if (lane < c) {
#pragma unroll
            for (int j = 0; j < CW; ++j)
                if (j < c) Rstack[leaf * c + lane + (long long)j * ldR] = (j >= lane) ? a[0][j] : 0.0;
        }
#pragma unroll
        for (int j = 0; j < CW; ++j)
            if (lane == j && j < c) tau_out[leaf * c + j] = tau[j];
    }
}

// B = H_1 ... H_c [W_leaf; 0]  for every leaf (W_leaf = rows leaf*c.. of W), written to Out (may alias V).
template <int CW, int RPL>
__global__ void __launch_bounds__(kTsqrThreads)
k_tsqr_apply(long long nrows, int c, const double* V, long long ldV, const double* __restrict__ tau_in,
             const double* __restrict__ W, long long ldW, double* Out, long long ldOut) {
    constexpr int LEAF = 32 * RPL;
    const int lane = threadIdx.x & 31;
    const long long nleaves = (nrows + LEAF - 1) / LEAF;
    const long long wstride = (long long)gridDim.x * (kTsqrThreads / 32);
    for (long long leaf = (long long)blockIdx.x * (kTsqrThreads / 32) + (threadIdx.x >> 5); leaf < nleaves; leaf += wstride) {
        const long long r0 = leaf * LEAF;
        double b[RPL][CW];
#pragma unroll
        for (int i = 0; i < RPL; ++i)
#pragma unroll
            for (int j = 0; j < CW; ++j)
                b[i][j] = (i == 0 && lane < c && j < c) ? W[leaf * c + lane + (long long)j * ldW] : 0.0;
#pragma unroll
        for (int jj = 0; jj < CW; ++jj) {
            const int j = CW - 1 - jj;
            if (j < c) {
                const double tj = tau_in[leaf * c + j];
                // reflector j: 1 at local row j, stored tail below, 0 above (streamed from V, L1/L2 resident)
                double v[RPL];
#pragma unroll
                for (int i = 0; i < RPL; ++i) {
                    const long long row = r0 + i * 32 + lane;
                    const int lrow = i * 32 + lane;
                    const double x = (row < nrows) ? V[row + (long long)j * ldV] : 0.0;
                    v[i] = (lrow > j) ? x : (lrow == j ? 1.0 : 0.0);
                }
                if (tj != 0.0) {
#pragma unroll
                    for (int l = 0; l < CW; ++l) {
                        if (l < c) {
                            double w = 0.0;
#pragma unroll
                            for (int i = 0; i < RPL; ++i) w = fma(v[i], b[i][l], w);
                            w = warp_sum(w) * tj;
#pragma unroll
                            for (int i = 0; i < RPL; ++i) b[i][l] = fma(-w, v[i], b[i][l]);
                        }
                    }
                }
            }
        }
#pragma unroll
        for (int i = 0; i < RPL; ++i) {
            const long long row = r0 + i * 32 + lane;
#pragma unroll
            for (int j = 0; j < CW; ++j)
                if (row < nrows && j < c) Out[row + (long long)j * ldOut] = b[i][j];
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

// place this rank's c x c R into its slot of a zeroed (P*c) x c stack (then all-reduced = all-gathered)
__global__ void k_tsqr_slot(int c, int P, int rank, const double* __restrict__ Rloc, long long ldR, double* __restrict__ S,
                            const int* __restrict__ pred, int want) {
    const bool run = !(pred && *pred != want);
    for (int e = threadIdx.x; e < P * c * c; e += blockDim.x) {
        const int row = e % (P * c), j = e / (P * c);
        const int q = row / c, i = row % c;
        // a predicated-off call still has to feed zeros/identical data into the collective
        S[e] = (run && q == rank) ? Rloc[i + (long long)j * ldR] : 0.0;
    }
}

struct Level {
    double* V; long long ldV; long long nrows; long long leaves; double* tau; double* Rstack; long long ldR;
};

struct Plan {
    std::vector<Level> levels;
    int c = 0, rpl = 0;
    long long n = 0;
    double* D = nullptr;       // c x c sign matrix handed to the top leaf
    double* Wtop = nullptr;    // what the local top leaf receives (multi-rank: slice of the global walk)
    long long ldWtop = 0;
    int first_global = -1;     // index of the first level that works on the gathered stack
};

// one TSQR factorisation is live per context at a time (calls are serialised)
std::map<calz_ctx*, Plan> g_plans;

template <int CW, int RPL>
int run_leaf(calz_ctx* ctx, const Level& L, int c, const double* A, long long ldA, const int* pred, int want) {
    const long long warps = (L.leaves + 0);
    constexpr int WPC = kTsqrThreads / 32;
    int grid = (int)std::min<long long>((warps + WPC - 1) / WPC, (long long)ctx->num_sms * 16);
    k_tsqr_leaf<CW, RPL><<<std::max(grid, 1), kTsqrThreads, 0, ctx->stream>>>(L.nrows, c, A, ldA, L.V, L.ldV, L.tau, L.Rstack, L.ldR, pred, want);
    CALZ_LAUNCH_CHECK(ctx);
    return CALZ_OK;
}

template <int CW, int RPL>
int run_apply(calz_ctx* ctx, const Level& L, int c, const double* W, long long ldW, double* Out, long long ldOut) {
    constexpr int WPC = kTsqrThreads / 32;
    int grid = (int)std::min<long long>((L.leaves + WPC - 1) / WPC, (long long)ctx->num_sms * 16);
    k_tsqr_apply<CW, RPL><<<std::max(grid, 1), kTsqrThreads, 0, ctx->stream>>>(L.nrows, c, L.V, L.ldV, L.tau, W, ldW, Out, ldOut);
    CALZ_LAUNCH_CHECK(ctx);
    return CALZ_OK;
}

int rpl_for(int c) { return c <= 8 ? 8 : (c <= 16 ? 4 : 2); }

int dispatch_leaf(calz_ctx* ctx, const Level& L, int c, const double* A, long long ldA, const int* pred, int want) {
    if (c <= 8) return run_leaf<8, 8>(ctx, L, c, A, ldA, pred, want);
    if (c <= 16) return run_leaf<16, 4>(ctx, L, c, A, ldA, pred, want);
    if (c <= 24) return run_leaf<24, 2>(ctx, L, c, A, ldA, pred, want);
    return run_leaf<32, 2>(ctx, L, c, A, ldA, pred, want);
}

int dispatch_apply(calz_ctx* ctx, const Level& L, int c, const double* W, long long ldW, double* Out, long long ldOut) {
    if (c <= 8) return run_apply<8, 8>(ctx, L, c, W, ldW, Out, ldOut);
    if (c <= 16) return run_apply<16, 4>(ctx, L, c, W, ldW, Out, ldOut);
    if (c <= 24) return run_apply<24, 2>(ctx, L, c, W, ldW, Out, ldOut);
    return run_apply<32, 2>(ctx, L, c, W, ldW, Out, ldOut);
}

}  // namespace

int tsqr_factor(calz_ctx* ctx, int64_t n, int c, const double* A, int64_t ldA, double* R_dev, const int* pred, int want) {
    if (c < 1 || c > kMaxC) return set_error(ctx, CALZ_ERR_UNSUPPORTED, "tsqr: c=%d outside [1,%d]", c, kMaxC);
    const int P = ctx->nranks;
    const int rpl = rpl_for(c);
    const long long LEAF = 32LL * rpl;
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
    if (P > 1) doubles += (size_t)P * c * c;
    const long long ldV0 = round_up(n, 32);
    CALZ_TRY(reserve(ctx, ctx->work[0], (size_t)ldV0 * c * sizeof(double)));
    CALZ_TRY(reserve(ctx, ctx->tsqr_r, doubles * sizeof(double)));
    double* p = (double*)ctx->tsqr_r.p;
    Plan& plan = g_plans[ctx];
    plan.levels.assign(nlev, Level{});
    plan.c = c; plan.rpl = rpl; plan.n = n; plan.first_global = first_global;
    plan.D = p; p += 2 * (size_t)c * c;
    double* gathered = nullptr;
    if (P > 1) { gathered = p; p += (size_t)P * c * c; }
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
        if (l == 0) { L.V = (double*)ctx->work[0].p; L.ldV = ldV0; }
        else if (l == first_global) { L.V = gathered; L.ldV = (long long)P * c; }
        else { L.V = plan.levels[l - 1].Rstack; L.ldV = plan.levels[l - 1].ldR; }     // reflectors overwrite the stack
    }
    // ---- bottom-up
    for (int l = 0; l < nlev; ++l) {
        const Level& L = plan.levels[l];
        if (l == first_global) {
            const Level& T = plan.levels[l - 1];
            k_tsqr_slot<<<1, 256, 0, ctx->stream>>>(c, P, ctx->rank, T.Rstack, T.ldR, gathered, pred, want);
            CALZ_LAUNCH_CHECK(ctx);
            CALZ_TRY(allreduce_sum(ctx, gathered, (size_t)P * c * c));
        }
        const double* src = (l == 0) ? A : L.V;
        const long long lds = (l == 0) ? ldA : L.ldV;
        CALZ_TRY(dispatch_leaf(ctx, L, c, src, lds, pred, want));
    }
    const Level& T = plan.levels[nlev - 1];
    k_tsqr_finish<<<1, 256, 0, ctx->stream>>>(c, T.Rstack, T.ldR, R_dev, plan.D, pred, want);
    CALZ_LAUNCH_CHECK(ctx);
    return CALZ_OK;
}

int tsqr_form_q(calz_ctx* ctx, int64_t n, int c, const double* A, int64_t ldA, double* Q, int64_t ldQ) {
    (void)A; (void)ldA;    // the reflectors of the last tsqr_factor that ran are on the device
    Plan& plan = g_plans[ctx];
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

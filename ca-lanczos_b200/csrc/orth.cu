// Block orthogonalisation entry points: tsqr / cholqr / normalize / project / projectAndNormalize.
//
// The control flow of projectAndNormalize.m:3-90 is kept exactly -- norms before, project, normalize, the 50 %
// norm-drop test, the second pass on the UN-normalised Y, the coefficient sum -- but it runs without a host
// round trip: the test result is a device flag and the pass-2 kernels are predicated on it, so one call is
// one uninterrupted kernel train (plus small all-reduces when row-partitioned) and a single D2H copy of
// the small results at the end.  Work the reference computes and then discards when pass 2 fires (the Q of
// the first QR, projectAndNormalize.m:26 vs :63-64) is simply not launched.
#include <string.h>

#include <algorithm>
#include <map>

#include "tsops.cuh"

using namespace calz;

namespace {

struct SmallLayout {            // device scratch for one projectAndNormalize call (doubles)
    int* flags;                 // [0] second pass, [1] chol info pass 1, [2] chol info pass 2
    std::vector<double*> C1, C2;
    std::vector<int> ld1;
    double *G, *R1, *R2;
    size_t doubles;
};

int small_scratch(calz_ctx* ctx, size_t doubles, double** p) {
    CALZ_TRY(reserve(ctx, ctx->small, doubles * sizeof(double)));
    if (doubles * sizeof(double) > ctx->pinned_bytes) {
        CALZ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        if (ctx->pinned) cudaFreeHost(ctx->pinned);
        ctx->pinned_bytes = doubles * sizeof(double) * 2;
        CALZ_CUDA(ctx, cudaMallocHost(&ctx->pinned, ctx->pinned_bytes));
    }
    *p = (double*)ctx->small.p;
    return CALZ_OK;
}

bool empty_block(const double* const* Qblk, const int* mcols, int i) { return !Qblk || !Qblk[i] || !mcols || mcols[i] <= 0; }

// QR of S (n x c) by `backend`: R -> R_dev; with CholQR the Gram matrix goes through G_dev.
// decision (optional): norm-drop flag from the columns of R against nb2.  cond_dev (CholQR2): "needs another pass".
int factor_r(calz_ctx* ctx, int backend, int64_t n, int c, const double* S, int64_t ldS, double* G_dev, double* R_dev,
             int* info_dev, const double* nb2, int nb2_stride, int* flag_dev, const int* pred, int want, int* cond_dev = nullptr) {
    if (backend == CALZ_QR_CHOLQR || backend == CALZ_QR_CHOLQR2) {
        CALZ_TRY(tsmm_tn(ctx, n, one_panel(S, ldS, c), S, ldS, c, G_dev, c, true, pred, want, true));
        return chol_small(ctx, c, G_dev, R_dev, info_dev, nb2, nb2_stride, flag_dev, pred, want, backend == CALZ_QR_CHOLQR2,
                          backend == CALZ_QR_CHOLQR2 ? cond_dev : nullptr);
    }
    CALZ_TRY(tsqr_factor(ctx, n, c, S, ldS, R_dev, pred, want));
    if (nb2) CALZ_TRY(norm_drop_decision(ctx, c, R_dev, nb2, nb2_stride, flag_dev));
    return CALZ_OK;
}

// CholQR2 refinement, driven from the host AFTER the small results came back (so the steady state, where the
// conditioning estimate is fine, pays nothing): repeat { G = Q'Q; Rb = chol(G); Q = Q/Rb; Rfin = Rb*Rfin } while the
// estimate still asks for it (at most 3 passes: a shifted first pass needs two more, "shifted CholeskyQR3").
// scratch: 2*c*c doubles + 2 ints on the device.  Rfin_dev is updated in place; returns the last chol info.
int cholqr_refine(calz_ctx* ctx, int64_t n, int c, double* Q, int64_t ldQ, double* Rfin_dev, double* scratch, int* ints_dev,
                  int* info_out) {
    const size_t cc = (size_t)c * c;
    double *G2 = scratch, *Rb = scratch + cc;
    *info_out = 0;
    for (int pass = 0; pass < 3; ++pass) {
        CALZ_TRY(tsmm_tn(ctx, n, one_panel(Q, ldQ, c), Q, ldQ, c, G2, c, true, nullptr, 0, true));
        CALZ_TRY(chol_small(ctx, c, G2, Rb, ints_dev, nullptr, 0, nullptr, nullptr, 0, true, ints_dev + 1));
        CALZ_TRY(ts_trsolve(ctx, n, c, Q, ldQ, Rb, Q, ldQ, nullptr, 0));
        CALZ_TRY(rmul_upper(ctx, c, Rb, Rfin_dev, nullptr, 0));
        int h[2];
        CALZ_CUDA(ctx, cudaMemcpyAsync(h, ints_dev, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
        CALZ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        if (h[0] > 0) { *info_out = h[0]; break; }
        if (!h[1]) break;
    }
    return CALZ_OK;
}

// ---- projections on the tile kernels for ANY block list (tiles.cu): a block of more than tile_panel_width(c) columns is cut
// into near-equal column panels; every panel owns a dense (m + c) x c result S = [Q_p Y]'Y per sweep.
struct PanelRef { int k, a0, m; size_t off1, off2; };         // k: position in the list of non-empty blocks
void split_panels(int k, int M, int W, std::vector<PanelRef>& out) {
    const int np = (M + W - 1) / W;
    for (int p = 0, a0 = 0; p < np; ++p) {
        const int m = (M - a0 + (np - p) - 1) / (np - p);
        out.push_back(PanelRef{k, a0, m, 0, 0});
        a0 += m;
    }
}

// One sweep of project.m:32-39, block after block: C_i = Q_i'Y -- one coefficient pass per panel, ALL of them before the block's
// first update (classical Gram-Schmidt inside a block, modified across blocks, like the reference) -- then Y = Y - Q_i*C_i panel
// by panel.  The very last update also contracts Y'Y into SG (rows [m_last, m_last + c), leading dimension m_last + c), so the
// norms-after / the Gram matrix of the following CholQR cost no further pass.  src -> dst on the first update, in place after it.
int panel_sweep(calz_ctx* ctx, int64_t n, const double* const* Qblk, const int64_t* ldQ, const std::vector<int>& blocks,
                const std::vector<PanelRef>& panels, int which, int c, const double*& src, int64_t& ldsrc, double* dst, int64_t lddst,
                double* sm, double* SG, const int* pred, int want) {
    size_t i = 0;
    while (i < panels.size()) {
        size_t j = i;
        while (j < panels.size() && panels[j].k == panels[i].k) ++j;
        const int b = blocks[panels[i].k];
        for (size_t p = i; p < j; ++p) {
            const PanelRef& P = panels[p];
            CALZ_TRY(tile_pass(ctx, 0, n, Qblk[b] + (int64_t)P.a0 * ldQ[b], ldQ[b], P.m, src, ldsrc, c, nullptr, 0, nullptr, 0,
                               sm + (which == 1 ? P.off1 : P.off2), P.m + c, pred, want, true));
        }
        for (size_t p = i; p < j; ++p) {
            const PanelRef& P = panels[p];
            const bool last = p + 1 == panels.size();
            CALZ_TRY(tile_pass(ctx, 2, n, Qblk[b] + (int64_t)P.a0 * ldQ[b], ldQ[b], P.m, src, ldsrc, c,
                               sm + (which == 1 ? P.off1 : P.off2), P.m + c, dst, lddst, last ? SG : nullptr, P.m + c, pred, want, true));
            src = dst;
            ldsrc = lddst;
        }
        i = j;
    }
    return CALZ_OK;
}

// cholqr.m (single pass) or its CholQR2 variant: R on the host, info = failing pivot (0: none)
int cholqr_device(calz_ctx* ctx, int64_t n, int c, const double* X, int64_t ldX, double* Q, int64_t ldQ, double* R,
                  int* info, bool adaptive) {
    const size_t cc = (size_t)c * c;
    double* sm;
    CALZ_TRY(small_scratch(ctx, 8 + 5 * cc, &sm));
    int* flags = (int*)sm;                       // [0] info, [1] cond, [2..3] refinement scratch
    CALZ_CUDA(ctx, cudaMemsetAsync(sm, 0, 8 * sizeof(double), ctx->stream));
    double *G = sm + 8, *R1 = G + cc, *scr = R1 + cc;
    CALZ_TRY(factor_r(ctx, adaptive ? CALZ_QR_CHOLQR2 : CALZ_QR_CHOLQR, n, c, X, ldX, G, R1, flags, nullptr, 0, nullptr, nullptr, 0,
                      flags + 1));
    CALZ_TRY(ts_trsolve(ctx, n, c, X, ldX, R1, Q, ldQ, nullptr, 0));
    CALZ_CUDA(ctx, cudaMemcpyAsync(ctx->pinned, sm, (8 + 2 * cc) * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CALZ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    const int* hf = (const int*)ctx->pinned;
    int inf = hf[0] > 0 ? hf[0] : 0;
    if (!inf && adaptive && hf[1]) {
        CALZ_TRY(cholqr_refine(ctx, n, c, Q, ldQ, R1, scr, flags + 2, &inf));
        CALZ_CUDA(ctx, cudaMemcpyAsync(ctx->pinned + 8 + cc, R1, cc * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        CALZ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    memcpy(R, ctx->pinned + 8 + cc, cc * sizeof(double));
    if (info) *info = inf;
    if (inf) return set_error(ctx, CALZ_ERR_CHOL, "cholqr: Gram matrix not positive definite at pivot %d", inf);
    return CALZ_OK;
}

}  // namespace

extern "C" {

int calz_gram(calz_ctx* ctx, int64_t n, int m, const double* A, int64_t ldA, int c, const double* B, int64_t ldB, double* C_dev) {
    if (!ctx || !A || !B || !C_dev || n < 0 || m < 1 || c < 1) return set_error(ctx, CALZ_ERR_BADARG, "calz_gram: bad arguments");
    CALZ_CUDA(ctx, cudaSetDevice(ctx->device));
    return tsmm_tn(ctx, n, one_panel(A, ldA, m), B, ldB, c, C_dev, m, A == B && ldA == ldB && m == c, nullptr, 0, true);
}

int calz_block_axpy(calz_ctx* ctx, int64_t n, int m, const double* Q, int64_t ldQ, int c, const double* C_host, const double* X,
                    int64_t ldX, double* Y, int64_t ldY) {
    if (!ctx || !Q || !C_host || !Y || n < 1 || m < 1 || c < 1 || c > kMaxC)
        return set_error(ctx, CALZ_ERR_BADARG, "calz_block_axpy: bad arguments");
    CALZ_CUDA(ctx, cudaSetDevice(ctx->device));
    double* sm;
    CALZ_TRY(small_scratch(ctx, (size_t)m * c, &sm));
    // the coefficient block travels through pinned staging so that the copy is stream-ordered and the call stays asynchronous
    CALZ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));          // the staging area may still be read by an earlier copy
    memcpy(ctx->pinned, C_host, (size_t)m * c * sizeof(double));
    CALZ_CUDA(ctx, cudaMemcpyAsync(sm, ctx->pinned, (size_t)m * c * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    return ts_update(ctx, n, Q, ldQ, m, sm, m, X, X ? ldX : 0, c, Y, ldY, nullptr, 0);
}

int calz_orth_error(calz_ctx* ctx, int64_t n, int nblk, const double* const* Qblk, const int64_t* ldQ, const int* mcols, int mode,
                    int s, double* err) {
    if (!ctx || !err || n < 1 || nblk < 1 || !Qblk || !ldQ || !mcols || (mode != CALZ_ORTH_FRO && mode != CALZ_ORTH_LASTBLOCK))
        return set_error(ctx, CALZ_ERR_BADARG, "calz_orth_error: bad arguments");
    CALZ_CUDA(ctx, cudaSetDevice(ctx->device));
    // virtual column j of Q -> (pointer, ld)
    std::vector<const double*> colp;
    for (int i = 0; i < nblk; ++i)
        if (!empty_block(Qblk, mcols, i))
            for (int j = 0; j < mcols[i]; ++j) colp.push_back(Qblk[i] + (size_t)j * ldQ[i]);
    const int tot = (int)colp.size();
    if (tot == 0) { *err = 0.0; return CALZ_OK; }
    // runs of columns that are contiguous in memory with one leading dimension (a block is one run)
    struct Run { const double* p; int64_t ld; int c0, nc; };
    std::vector<Run> runs;
    {
        int c0 = 0;
        for (int i = 0; i < nblk; ++i)
            if (!empty_block(Qblk, mcols, i)) { runs.push_back({Qblk[i], ldQ[i], c0, mcols[i]}); c0 += mcols[i]; }
    }
    // G(:, cols of B-chunk) = Q' * B-chunk, B-chunks of <= 32 columns inside one run; only the column range the mode needs
    const int jlo = (mode == CALZ_ORTH_LASTBLOCK && tot > s + 1) ? tot - s - 1 : 0;
    std::vector<double> G((size_t)tot * tot, 0.0);
    double* sm;
    CALZ_TRY(small_scratch(ctx, (size_t)tot * 32, &sm));
    for (const Run& rb : runs) {
        for (int b0 = 0; b0 < rb.nc; b0 += 32) {
            const int cb = std::min(32, rb.nc - b0);
            if (rb.c0 + b0 + cb <= jlo) continue;
            const double* B = rb.p + (size_t)b0 * rb.ld;
            // A side: up to 4 runs per launch (Panels), rows of G at the run's column offset
            for (size_t r0 = 0; r0 < runs.size(); r0 += 4) {
                Panels A{};
                int a_c0 = runs[r0].c0;
                for (size_t r = r0; r < std::min(runs.size(), r0 + 4); ++r) add_panel(A, runs[r].p, runs[r].ld, runs[r].nc);
                CALZ_TRY(tsmm_tn(ctx, n, A, B, rb.ld, cb, sm, A.total, false, nullptr, 0, true));
                CALZ_CUDA(ctx, cudaMemcpyAsync(ctx->pinned, sm, (size_t)A.total * cb * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
                CALZ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
                for (int j = 0; j < cb; ++j)
                    for (int a = 0; a < A.total; ++a) G[(size_t)(rb.c0 + b0 + j) * tot + a_c0 + a] = ctx->pinned[(size_t)j * A.total + a];
            }
        }
    }
    double e = 0.0;
    if (mode == CALZ_ORTH_FRO) {
        for (int j = 0; j < tot; ++j)
            for (int i = 0; i < tot; ++i) {
                const double d = (i == j ? 1.0 : 0.0) - G[(size_t)j * tot + i];
                e += d * d;
            }
        e = sqrt(e);
    } else if (tot > s + 1) {
        for (int j = tot - s - 1; j < tot; ++j)
            for (int i = 0; i < tot - s - 1; ++i) e = std::max(e, fabs(G[(size_t)j * tot + i]));
    } else {
        for (int j = 0; j < tot; ++j)
            for (int i = 0; i < tot; ++i) e = std::max(e, fabs(G[(size_t)j * tot + i] - (i == j ? 1.0 : 0.0)));
    }
    *err = e;
    return CALZ_OK;
}

int calz_cholqr(calz_ctx* ctx, int64_t n, int c, const double* X, int64_t ldX, double* Q, int64_t ldQ, double* R, int* info) {
    if (!ctx || !X || !Q || !R || n < 1 || c < 1 || c > kMaxC) return set_error(ctx, CALZ_ERR_BADARG, "calz_cholqr: bad arguments");
    CALZ_CUDA(ctx, cudaSetDevice(ctx->device));
    return cholqr_device(ctx, n, c, X, ldX, Q, ldQ, R, info, false);      // cholqr.m:3-8 as written: one pass
}

int calz_tsqr(calz_ctx* ctx, int64_t n, int c, const double* A, int64_t ldA, double* Q, int64_t ldQ, double* R) {
    if (!ctx || !A || !Q || !R || n < 1 || c < 1 || c > kMaxC) return set_error(ctx, CALZ_ERR_BADARG, "calz_tsqr: bad arguments");
    CALZ_CUDA(ctx, cudaSetDevice(ctx->device));
    double* sm;
    CALZ_TRY(small_scratch(ctx, (size_t)c * c, &sm));
    CALZ_TRY(tsqr_factor(ctx, n, c, A, ldA, sm, nullptr, 0));
    CALZ_TRY(tsqr_form_q(ctx, n, c, A, ldA, Q, ldQ));
    CALZ_CUDA(ctx, cudaMemcpyAsync(ctx->pinned, sm, (size_t)c * c * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CALZ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    memcpy(R, ctx->pinned, (size_t)c * c * sizeof(double));
    return CALZ_OK;
}

int calz_normalize(calz_ctx* ctx, int64_t n, int c, const double* X, int64_t ldX, int backend, double tol, double* Q,
                   int64_t ldQ, double* R, int* rank) {
    if (tol <= 0) tol = 1.0e-8;      // normalize.m:8-10
    int st;
    if (backend == CALZ_QR_CHOLQR || backend == CALZ_QR_CHOLQR2) {
        if (!ctx || !X || !Q || !R || n < 1 || c < 1 || c > kMaxC) return set_error(ctx, CALZ_ERR_BADARG, "calz_normalize: bad arguments");
        CALZ_CUDA(ctx, cudaSetDevice(ctx->device));
        int info = 0;
        st = cholqr_device(ctx, n, c, X, ldX, Q, ldQ, R, &info, backend == CALZ_QR_CHOLQR2);
    } else if (backend == CALZ_QR_TSQR) {
        st = calz_tsqr(ctx, n, c, X, ldX, Q, ldQ, R);
    } else {
        return set_error(ctx, CALZ_ERR_BADARG, "calz_normalize: unknown backend %d", backend);
    }
    CALZ_TRY(st);
    if (rank) *rank = numerical_rank(c, R, c, tol);
    return CALZ_OK;
}

int calz_project(calz_ctx* ctx, int64_t n, int nblk, const double* const* Qblk, const int64_t* ldQ, const int* mcols,
                 int c, double* X, int64_t ldX, int doreorth, double* const* Rblk) {
    if (!ctx || !X || n < 1 || c < 1 || c > kMaxC || nblk < 0) return set_error(ctx, CALZ_ERR_BADARG, "calz_project: bad arguments");
    CALZ_CUDA(ctx, cudaSetDevice(ctx->device));
    std::vector<int> blocks;
    for (int i = 0; i < nblk; ++i)
        if (!empty_block(Qblk, mcols, i)) blocks.push_back(i);
    if (!blocks.empty() && ctx->opt_tile_pipeline && ctx->opt_tile_panels && tile_panel_width(c) > 0) {
        // ---- tile kernels, panel by panel (shape-only condition: every rank takes the same path).  ||x_i||^2 before comes with the
        //      first coefficient pass ([Q_p X]'X), the norms after with the last update: no separate Gram passes (project.m:29-31,:41)
        std::vector<PanelRef> panels;
        for (size_t k = 0; k < blocks.size(); ++k) split_panels((int)k, mcols[blocks[k]], tile_panel_width(c), panels);
        size_t doubles = 0;
        for (PanelRef& P : panels) { P.off1 = doubles; doubles += (size_t)(P.m + c) * c; }
        for (PanelRef& P : panels) { P.off2 = doubles; doubles += (size_t)(P.m + c) * c; }
        const int m_last = panels.back().m, ldG = m_last + c;
        const size_t offSG = doubles; doubles += (size_t)ldG * c;
        double* sm;
        CALZ_TRY(small_scratch(ctx, doubles, &sm));
        CALZ_CUDA(ctx, cudaMemsetAsync(sm, 0, doubles * sizeof(double), ctx->stream));
        const double* src = X;
        int64_t ldsrc = ldX;
        CALZ_TRY(panel_sweep(ctx, n, Qblk, ldQ, blocks, panels, 1, c, src, ldsrc, X, ldX, sm, doreorth ? sm + offSG : nullptr, nullptr, 0));
        bool second = false;
        if (doreorth) {
            CALZ_CUDA(ctx, cudaMemcpyAsync(ctx->pinned, sm, doubles * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
            CALZ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
            const int ld0 = panels[0].m + c;
            double worst = -1e300;
            for (int j = 0; j < c; ++j) {
                const double b = sqrt(ctx->pinned[panels[0].off1 + (size_t)j * ld0 + panels[0].m + j]);
                const double a = sqrt(ctx->pinned[offSG + (size_t)j * ldG + m_last + j]);
                worst = std::max(worst, 0.5 * b - a);
            }
            second = worst < 0;             // project.m:43-46, criterion kept as written
            if (second) CALZ_TRY(panel_sweep(ctx, n, Qblk, ldQ, blocks, panels, 2, c, src, ldsrc, X, ldX, sm, nullptr, nullptr, 0));
        }
        CALZ_CUDA(ctx, cudaMemcpyAsync(ctx->pinned, sm, doubles * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        CALZ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        for (const PanelRef& P : panels) {
            const int i = blocks[P.k], m = mcols[i], ld = P.m + c;
            if (!Rblk || !Rblk[i]) continue;
            for (int j = 0; j < c; ++j)
                for (int a = 0; a < P.m; ++a)
                    Rblk[i][(size_t)j * m + P.a0 + a] = ctx->pinned[P.off1 + (size_t)j * ld + a] + (second ? ctx->pinned[P.off2 + (size_t)j * ld + a] : 0.0);
        }
        return CALZ_OK;
    }
    size_t tot = 0;
    for (int i = 0; i < nblk; ++i)
        if (!empty_block(Qblk, mcols, i)) tot += (size_t)mcols[i] * c;
    double* sm;
    const size_t need = 2 * tot + 2 * (size_t)c * c + 2 * c;
    CALZ_TRY(small_scratch(ctx, need, &sm));
    double *C1 = sm, *C2 = sm + tot, *nbef = C2 + tot, *naft = nbef + (size_t)c * c;
    const int sweeps_max = doreorth ? 2 : 1;
    bool second = false;
    if (doreorth) {   // project.m:29-31 -- column norms before (diag of X'X)
        CALZ_TRY(tsmm_tn(ctx, n, one_panel(X, ldX, c), X, ldX, c, nbef, c, true, nullptr, 0, true));
    }
    for (int sweep = 0; sweep < sweeps_max; ++sweep) {
        double* C = sweep == 0 ? C1 : C2;
        size_t off = 0;
        for (int i = 0; i < nblk; ++i) {
            if (empty_block(Qblk, mcols, i)) continue;
            const int m = mcols[i];
            CALZ_TRY(tsmm_tn(ctx, n, one_panel(Qblk[i], ldQ[i], m), X, ldX, c, C + off, m, false, nullptr, 0, true));
            CALZ_TRY(ts_update(ctx, n, Qblk[i], ldQ[i], m, C + off, m, X, ldX, c, X, ldX, nullptr, 0));
            off += (size_t)m * c;
        }
        if (!doreorth || sweep == 1) break;
        // project.m:43-46 -- reorthogonalise when max(0.5*before - after) < 0  (criterion kept as written)
        CALZ_TRY(tsmm_tn(ctx, n, one_panel(X, ldX, c), X, ldX, c, naft, c, true, nullptr, 0, true));
        CALZ_CUDA(ctx, cudaMemcpyAsync(ctx->pinned, nbef, (2 * (size_t)c * c) * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        CALZ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        double worst = -1e300;
        for (int j = 0; j < c; ++j) {
            const double b = sqrt(ctx->pinned[(size_t)j * c + j]), a = sqrt(ctx->pinned[(size_t)c * c + (size_t)j * c + j]);
            worst = std::max(worst, 0.5 * b - a);
        }
        second = worst < 0;
        if (!second) break;
    }
    CALZ_CUDA(ctx, cudaMemcpyAsync(ctx->pinned, sm, 2 * tot * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CALZ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    size_t off = 0;
    for (int i = 0; i < nblk; ++i) {
        if (empty_block(Qblk, mcols, i)) continue;
        const size_t cnt = (size_t)mcols[i] * c;
        if (Rblk && Rblk[i])
            for (size_t e = 0; e < cnt; ++e) Rblk[i][e] = ctx->pinned[off + e] + (second ? ctx->pinned[tot + off + e] : 0.0);
        off += cnt;
    }
    return CALZ_OK;
}

// ---- one projectAndNormalize call in flight: where its small results live and how to unpack them
struct PanSlot {
    bool busy = false;
    cudaEvent_t ev = nullptr;
    double* host = nullptr;          // pinned
    size_t host_doubles = 0;
    int nb = 0, c = 0, backend = 0, nblk = 0;
    int64_t n = 0, ldQZ = 0;
    double* QZ = nullptr;
    double* sm = nullptr;            // device scratch base (valid until the next call on the stream reuses it)
    std::vector<int> blocks, mc;
    // where the coefficient pieces live in the scratch: panel p covers rows [a0, a0+m) of block pk[p] (leading dimensions pld1/pld2)
    std::vector<int> pk, pa0, pm, pld1, pld2;
    std::vector<size_t> poff1, poff2;
    size_t offRf = 0, offG2 = 0, doubles = 0;
};
constexpr int kPanSlots = 8;
struct PanRing { PanSlot slot[kPanSlots]; int next = 0; };

static int pan_enqueue(calz_ctx* ctx, int64_t n, int nblk, const double* const* Qblk, const int64_t* ldQ,
                       const int* mcols, int c, const double* X, int64_t ldX, int doreorth, int backend,
                       double* QZ, int64_t ldQZ, PanSlot& slot) {
    if (!ctx || !X || !QZ || n < 1 || c < 1 || c > kMaxC || nblk < 0 ||
        (backend != CALZ_QR_TSQR && backend != CALZ_QR_CHOLQR && backend != CALZ_QR_CHOLQR2))
        return set_error(ctx, CALZ_ERR_BADARG, "calz_project_and_normalize: bad arguments");
    if (QZ == X) return set_error(ctx, CALZ_ERR_BADARG, "calz_project_and_normalize: QZ must not alias X");
    CALZ_CUDA(ctx, cudaSetDevice(ctx->device));

    // ---- small device scratch
    std::vector<int> blocks;
    for (int i = 0; i < nblk; ++i)
        if (!empty_block(Qblk, mcols, i)) blocks.push_back(i);
    const int nb = (int)blocks.size();
    const bool fuse_norms = doreorth && nb > 0;        // column norms of X ride along with the first coefficient sweep
    // fused tile pipeline (tiles.cu): one previous block, Cholesky-based QR, TMA-compatible alignment
    const bool tile_shape = nb == 1 && fuse_norms && ctx->opt_tile_pipeline && mcols[blocks[0]] <= 16 && c <= 16;   // same on all ranks
    const bool tiled = tile_shape && tile_path_ok(n, Qblk[blocks[0]], ldQ[blocks[0]], mcols[blocks[0]], X, ldX, c, QZ, ldQZ);
    // tile_path_ok depends on the block shape only (rank-local alignment / row count are handled inside the kernels), so every
    // rank takes the same kernel path and issues the same sequence of collectives
    const bool fused_solve = tiled && backend != CALZ_QR_TSQR && ctx->opt_pan_fused_solve;
    // every other projection with the norm-drop test (several blocks -- {Qprev, Q_conv} of the restarted driver -- or one wide
    // block -- 'full' re-orthogonalisation): the same tile kernels, panel by panel.  Shape-only condition, same on all ranks.
    const bool panelled = !tile_shape && fuse_norms && ctx->opt_tile_pipeline && ctx->opt_tile_panels && tile_panel_width(c) > 0;
    size_t doubles = 8;                                // flags
    std::vector<size_t> off1(nb), off2(nb);
    std::vector<int> ld1(nb), ld2(nb);
    std::vector<PanelRef> panels;
    if (panelled) {
        for (int k = 0; k < nb; ++k) split_panels(k, mcols[blocks[k]], tile_panel_width(c), panels);
        for (PanelRef& P : panels) { P.off1 = doubles; doubles += (size_t)(P.m + c) * c; }
        for (PanelRef& P : panels) { P.off2 = doubles; doubles += (size_t)(P.m + c) * c; }
    } else {
        for (int k = 0; k < nb; ++k) {
            const int m = mcols[blocks[k]];
            ld1[k] = (k == 0 && fuse_norms) ? m + c : m;
            off1[k] = doubles; doubles += (size_t)ld1[k] * c;
        }
        for (int k = 0; k < nb; ++k) {
            ld2[k] = tiled ? mcols[blocks[k]] + c : mcols[blocks[k]];
            off2[k] = doubles; doubles += (size_t)ld2[k] * c;
        }
    }
    const int m_last = panelled ? panels.back().m : 0;
    const size_t offSG1 = doubles; doubles += panelled ? (size_t)(m_last + c) * c : 0;     // Gram of Y / of Z from the last update
    const size_t offSG2 = doubles; doubles += panelled ? (size_t)(m_last + c) * c : 0;
    const size_t offS3 = doubles; doubles += tiled ? (size_t)(mcols[blocks[0]] + c) * c : 0;
    const size_t offG = doubles; doubles += (size_t)c * c;
    const size_t offR1 = doubles; doubles += (size_t)c * c;
    const size_t offR2 = doubles; doubles += (size_t)c * c;
    const size_t offRf = doubles; doubles += (size_t)c * c;       // R of the last normalize (what the host reads)
    const size_t offG2 = doubles; doubles += 2 * (size_t)c * c;   // CholQR2 refinement scratch
    double* sm;
    CALZ_TRY(small_scratch(ctx, doubles, &sm));
    int* flags = (int*)sm;
    CALZ_CUDA(ctx, cudaMemsetAsync(sm, 0, doubles * sizeof(double), ctx->stream));
    double *G = sm + offG, *R1 = sm + offR1, *R2 = sm + offR2, *Rf = sm + offRf;
    // flags: [0] second pass, [1]/[2] chol info pass 1/2, [3]/[4] conditioning flag pass 1/2, [5] selected, [6..7] refinement

    const double* src = X;
    int64_t ldsrc = ldX;
    if (tiled) {
        const int i = blocks[0], m = mcols[i], ldS = m + c;
        double *S1 = sm + off1[0], *S2 = sm + off2[0], *S3 = sm + offS3;
        // pass 1: S1 = [Q X]'X  -> C1 and ||x_i||^2
        CALZ_TRY(tile_pass(ctx, 0, n, Qblk[i], ldQ[i], m, X, ldX, c, nullptr, 0, nullptr, 0, S1, ldS, nullptr, 0, true));
        // pass 2: Y = X - Q*C1 (into QZ), S2 = [Q Y]'Y -> C2 (used only if pass 2 of the reference fires) and G_Y.
        // Cholesky back ends (fused_solve): Y stays in registers / shared memory, it is never written to HBM
        CALZ_TRY(tile_pass(ctx, 1, n, Qblk[i], ldQ[i], m, X, ldX, c, S1, ldS, fused_solve ? nullptr : QZ, ldQZ, S2, ldS, nullptr, 0, true));
        if (fused_solve) {
            // R1, the norm-drop test, R2 from the downdated Gram and Rf in one launch; then ONE more pass over [Q | X]:
            // QZ = ((X - Q*C1) - Q*C2)/Rf if the test fired (pass 3 and the triangular solve fused), else QZ = (X - Q*C1)/Rf
            CALZ_TRY(chol_pan(ctx, c, m, S2, ldS, S1 + m, ldS + 1, backend == CALZ_QR_CHOLQR2, R1, R2, Rf, flags));
            CALZ_TRY(tile_update_solve(ctx, n, Qblk[i], ldQ[i], m, X, ldX, c, S1, ldS, S2, ldS, flags, Rf, QZ, ldQZ));
        } else {
            if (backend != CALZ_QR_TSQR)
                CALZ_TRY(chol_small(ctx, c, S2 + m, R1, flags + 1, S1 + m, ldS + 1, flags + 0, nullptr, 0, backend == CALZ_QR_CHOLQR2,
                                    backend == CALZ_QR_CHOLQR2 ? flags + 3 : nullptr, ldS));
            else   // TSQR: the norms-after of the test are sqrt(diag(Y'Y)); Y itself is only factorised if pass 2 does not fire
                CALZ_TRY(norm_drop_from_gram(ctx, c, S2 + m, ldS, S1 + m, ldS + 1, flags + 0));
            // pass 3 (iff the norm-drop test fired): Z = Y - Q*C2 in place, S3 = Z'Z
            CALZ_TRY(tile_pass(ctx, 2, n, Qblk[i], ldQ[i], m, QZ, ldQZ, c, S2, ldS, QZ, ldQZ, S3, ldS, flags, 1, true));
            if (backend != CALZ_QR_TSQR)
                CALZ_TRY(chol_small(ctx, c, S3 + m, R2, flags + 2, nullptr, 0, nullptr, flags, 1, backend == CALZ_QR_CHOLQR2,
                                    backend == CALZ_QR_CHOLQR2 ? flags + 4 : nullptr, ldS));
            else   // one Householder TSQR of whatever QZ holds now (Z, or Y if pass 2 did not fire): the LAST normalize
                CALZ_TRY(tsqr_factor(ctx, n, c, QZ, ldQZ, R1, nullptr, 0));
        }
        src = QZ;
        ldsrc = ldQZ;
    } else if (panelled) {
        const int ldG = m_last + c;
        double *SG1 = sm + offSG1, *SG2 = sm + offSG2;
        // ---- pass 1: Y = X - sum_i Q_i (Q_i'X) into QZ; ||x_i||^2 = diag of the X'X part of the first panel's S; G_Y from the last update
        CALZ_TRY(panel_sweep(ctx, n, Qblk, ldQ, blocks, panels, 1, c, src, ldsrc, QZ, ldQZ, sm, SG1, nullptr, 0));
        const double* nb2 = sm + panels[0].off1 + panels[0].m;
        const int nb2_stride = panels[0].m + c + 1;
        if (backend != CALZ_QR_TSQR)
            CALZ_TRY(chol_small(ctx, c, SG1 + m_last, R1, flags + 1, nb2, nb2_stride, flags + 0, nullptr, 0, backend == CALZ_QR_CHOLQR2,
                                backend == CALZ_QR_CHOLQR2 ? flags + 3 : nullptr, ldG));
        else
            CALZ_TRY(norm_drop_from_gram(ctx, c, SG1 + m_last, ldG, nb2, nb2_stride, flags + 0));
        // ---- pass 2 (iff the norm-drop test fired, projectAndNormalize.m:61-73): Z = Y - sum_i Q_i (Q_i'Y) in place, G_Z
        CALZ_TRY(panel_sweep(ctx, n, Qblk, ldQ, blocks, panels, 2, c, src, ldsrc, QZ, ldQZ, sm, SG2, flags, 1));
        if (backend != CALZ_QR_TSQR)
            CALZ_TRY(chol_small(ctx, c, SG2 + m_last, R2, flags + 2, nullptr, 0, nullptr, flags, 1, backend == CALZ_QR_CHOLQR2,
                                backend == CALZ_QR_CHOLQR2 ? flags + 4 : nullptr, ldG));
        else   // one Householder TSQR of whatever QZ holds now (Z, or Y if pass 2 did not fire): the LAST normalize
            CALZ_TRY(tsqr_factor(ctx, n, c, QZ, ldQZ, R1, nullptr, 0));
    } else {
        // ---- pass 1: Y = X - sum_i Q_i (Q_i' X)   (sequential over blocks, project.m:32-39); Y lives in QZ
        for (int k = 0; k < nb; ++k) {
            const int i = blocks[k], m = mcols[i];
            Panels A = one_panel(Qblk[i], ldQ[i], m);
            if (k == 0 && fuse_norms) add_panel(A, X, ldX, c);          // [Q_1 X]' X: last c rows hold X'X
            double* C = sm + off1[k];
            CALZ_TRY(tsmm_tn(ctx, n, A, src, ldsrc, c, C, ld1[k], false, nullptr, 0, true));
            CALZ_TRY(ts_update(ctx, n, Qblk[i], ldQ[i], m, C, ld1[k], src, ldsrc, c, QZ, ldQZ, nullptr, 0));
            src = QZ;
            ldsrc = ldQZ;
        }
        // nb2_i = ||X(:,i)||^2: diagonal of the X'X block of the fused sweep
        const double* nb2 = nullptr;
        int nb2_stride = 0;
        if (fuse_norms) {
            nb2 = sm + off1[0] + mcols[blocks[0]];
            nb2_stride = ld1[0] + 1;
        }
        // ---- normalize(Y): R1 (+ the norm-drop decision on the device)
        CALZ_TRY(factor_r(ctx, backend, n, c, src, ldsrc, G, R1, flags + 1, nb2, nb2_stride, flags + 0, nullptr, 0, flags + 3));

        // ---- pass 2, predicated on flags[0] == 1 (projectAndNormalize.m:61-73): Z = Y - sum_i Q_i (Q_i' Y), in place
        if (fuse_norms) {
            for (int k = 0; k < nb; ++k) {
                const int i = blocks[k], m = mcols[i];
                double* C = sm + off2[k];
                CALZ_TRY(tsmm_tn(ctx, n, one_panel(Qblk[i], ldQ[i], m), QZ, ldQZ, c, C, m, false, flags, 1, true));
                CALZ_TRY(ts_update(ctx, n, Qblk[i], ldQ[i], m, C, m, QZ, ldQZ, c, QZ, ldQZ, flags, 1));
            }
            CALZ_TRY(factor_r(ctx, backend, n, c, QZ, ldQZ, G, R2, flags + 2, nullptr, 0, nullptr, flags, 1, flags + 4));
        }
    }
    // ---- Q of the LAST normalize only
    if (fused_solve) {
        // done above
    } else if (backend != CALZ_QR_TSQR) {
        CALZ_TRY(select_r(ctx, c, R1, R2, flags, flags + 3, flags + 4, Rf, flags + 5));
        CALZ_TRY(ts_trsolve(ctx, n, c, src, ldsrc, Rf, QZ, ldQZ, nullptr, 0));
    } else {
        CALZ_TRY(select_r(ctx, c, R1, (tiled || panelled) ? R1 : R2, flags, nullptr, nullptr, Rf, nullptr));
        // the reflectors on the device belong to the last factorisation that actually ran (Y, or Z if pass 2 fired)
        CALZ_TRY(tsqr_form_q(ctx, n, c, src, ldsrc, QZ, ldQZ));
    }

    // ---- small results: one D2H copy into this call's pinned slot, completion event
    if (slot.host_doubles < doubles) {
        if (slot.host) cudaFreeHost(slot.host);
        slot.host_doubles = std::max<size_t>(doubles * 2, 4096);
        CALZ_CUDA(ctx, cudaMallocHost(&slot.host, slot.host_doubles * sizeof(double)));
    }
    if (!slot.ev) CALZ_CUDA(ctx, cudaEventCreateWithFlags(&slot.ev, cudaEventDisableTiming));
    CALZ_CUDA(ctx, cudaMemcpyAsync(slot.host, sm, doubles * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CALZ_CUDA(ctx, cudaEventRecord(slot.ev, ctx->stream));
    slot.busy = true;
    slot.nb = nb; slot.c = c; slot.backend = backend; slot.nblk = nblk; slot.n = n; slot.QZ = QZ; slot.ldQZ = ldQZ; slot.sm = sm;
    slot.blocks = blocks;
    slot.pk.clear(); slot.pa0.clear(); slot.pm.clear(); slot.pld1.clear(); slot.pld2.clear(); slot.poff1.clear(); slot.poff2.clear();
    if (panelled) {
        for (const PanelRef& P : panels) {
            slot.pk.push_back(P.k); slot.pa0.push_back(P.a0); slot.pm.push_back(P.m);
            slot.pld1.push_back(P.m + c); slot.pld2.push_back(P.m + c); slot.poff1.push_back(P.off1); slot.poff2.push_back(P.off2);
        }
    } else {
        for (int k = 0; k < nb; ++k) {
            slot.pk.push_back(k); slot.pa0.push_back(0); slot.pm.push_back(mcols[blocks[k]]);
            slot.pld1.push_back(ld1[k]); slot.pld2.push_back(ld2[k]); slot.poff1.push_back(off1[k]); slot.poff2.push_back(off2[k]);
        }
    }
    slot.mc.assign(nblk, 0);
    for (int i = 0; i < nblk; ++i) slot.mc[i] = empty_block(Qblk, mcols, i) ? 0 : mcols[i];
    slot.offRf = offRf; slot.offG2 = offG2; slot.doubles = doubles;
    return CALZ_OK;
}

// wait for the call, unpack the small results; allow_refine: run the (rare) CholQR2 refinement passes here
static int pan_finish(calz_ctx* ctx, PanSlot& slot, double* const* Rblk, double* Rlast, int* second_pass, int* rank,
                      bool allow_refine, int* needs_refine) {
    if (!slot.busy) return set_error(ctx, CALZ_ERR_BADARG, "projectAndNormalize: nothing to collect");
    CALZ_CUDA(ctx, cudaEventSynchronize(slot.ev));
    slot.busy = false;
    const int c = slot.c, nb = slot.nb, backend = slot.backend;
    const double* h = slot.host;
    const int* hf = (const int*)h;
    const bool second = hf[0] != 0;
    int info = second ? hf[2] : hf[1];
    if (info < 0) info = 0;                                              // negative = shifted retries (CholQR2), not a failure
    if (needs_refine) *needs_refine = 0;
    if (!info && backend == CALZ_QR_CHOLQR2 && hf[5]) {
        // rare path: the conditioning estimate asks for re-orthogonalisation (first block of an ill-conditioned start)
        if (!allow_refine) {
            if (needs_refine) *needs_refine = 1;
        } else {
            double* Rf = slot.sm + slot.offRf;
            CALZ_TRY(cholqr_refine(ctx, slot.n, c, slot.QZ, slot.ldQZ, Rf, slot.sm + slot.offG2, (int*)slot.sm + 6, &info));
            CALZ_CUDA(ctx, cudaMemcpyAsync(slot.host + slot.offRf, Rf, (size_t)c * c * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
            CALZ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        }
    }
    if (second_pass) *second_pass = second ? 1 : 0;
    (void)nb;
    for (size_t p = 0; p < slot.pk.size(); ++p) {
        const int i = slot.blocks[slot.pk[p]], m = slot.mc[i], a0 = slot.pa0[p];
        if (!Rblk || !Rblk[i]) continue;
        for (int j = 0; j < c; ++j)
            for (int a = 0; a < slot.pm[p]; ++a) {
                double v = h[slot.poff1[p] + (size_t)j * slot.pld1[p] + a];
                if (second) v = h[slot.poff2[p] + (size_t)j * slot.pld2[p] + a] + v;    // RZ{i} = RZ{i} + RY{i}  (:71-73)
                Rblk[i][(size_t)j * m + a0 + a] = v;
            }
    }
    const double* Rl = h + slot.offRf;
    if (Rlast) memcpy(Rlast, Rl, (size_t)c * c * sizeof(double));
    if (rank) *rank = numerical_rank(c, Rl, c, 1.0e-8);
    if (info) return set_error(ctx, CALZ_ERR_CHOL, "projectAndNormalize/cholqr: Gram matrix not positive definite at pivot %d (pass %d)",
                               info, second ? 2 : 1);
    return CALZ_OK;
}

static PanRing* ring_of(calz_ctx* ctx) {      // lives in the context (freed by calz_finalize through pan_ring_free)
    if (!ctx->pan_ring) ctx->pan_ring = new PanRing();
    return (PanRing*)ctx->pan_ring;
}

int calz_project_and_normalize(calz_ctx* ctx, int64_t n, int nblk, const double* const* Qblk, const int64_t* ldQ,
                               const int* mcols, int c, const double* X, int64_t ldX, int doreorth, int backend,
                               double* QZ, int64_t ldQZ, double* const* Rblk, double* Rlast, int* second_pass, int* rank) {
    int ticket = -1;
    CALZ_TRY(calz_project_and_normalize_async(ctx, n, nblk, Qblk, ldQ, mcols, c, X, ldX, doreorth, backend, QZ, ldQZ, &ticket));
    return pan_finish(ctx, ring_of(ctx)->slot[ticket], Rblk, Rlast, second_pass, rank, true, nullptr);
}

int calz_project_and_normalize_async(calz_ctx* ctx, int64_t n, int nblk, const double* const* Qblk, const int64_t* ldQ,
                                     const int* mcols, int c, const double* X, int64_t ldX, int doreorth, int backend,
                                     double* QZ, int64_t ldQZ, int* ticket) {
    if (!ctx || !ticket) return set_error(ctx, CALZ_ERR_BADARG, "calz_project_and_normalize_async: bad arguments");
    PanRing* ring = ring_of(ctx);
    const int t = ring->next;
    if (ring->slot[t].busy)
        return set_error(ctx, CALZ_ERR_BADARG, "projectAndNormalize: %d calls in flight, collect some first", kPanSlots);
    CALZ_TRY(pan_enqueue(ctx, n, nblk, Qblk, ldQ, mcols, c, X, ldX, doreorth, backend, QZ, ldQZ, ring->slot[t]));
    ring->next = (t + 1) % kPanSlots;
    *ticket = t;
    return CALZ_OK;
}

int calz_pan_collect(calz_ctx* ctx, int ticket, double* const* Rblk, double* Rlast, int* second_pass, int* rank, int* needs_refine) {
    if (!ctx || ticket < 0 || ticket >= kPanSlots) return set_error(ctx, CALZ_ERR_BADARG, "calz_pan_collect: bad ticket");
    return pan_finish(ctx, ring_of(ctx)->slot[ticket], Rblk, Rlast, second_pass, rank, false, needs_refine);
}

}  // extern "C"

namespace calz {
void pan_ring_free(calz_ctx* ctx) {
    PanRing* r = (PanRing*)ctx->pan_ring;
    if (!r) return;
    for (PanSlot& s : r->slot) {
        if (s.ev) cudaEventDestroy(s.ev);
        if (s.host) cudaFreeHost(s.host);
    }
    delete r;
    ctx->pan_ring = nullptr;
}
}  // namespace calz

extern "C" {

// ------------------------------------------------------------------------------------------ host flavours
namespace {
struct Staged {
    double* d = nullptr;
    int64_t ld = 0;
};
int stage_in(calz_ctx* ctx, DevBuf& buf, int64_t n, int cols, const double* h, int64_t ldh, Staged* out, bool copy) {
    out->ld = round_up(n, 32);
    CALZ_TRY(reserve(ctx, buf, (size_t)out->ld * std::max(cols, 1) * sizeof(double)));
    out->d = (double*)buf.p;
    if (copy && cols > 0)
        CALZ_CUDA(ctx, cudaMemcpy2DAsync(out->d, (size_t)out->ld * sizeof(double), h, (size_t)ldh * sizeof(double),
                                         (size_t)n * sizeof(double), (size_t)cols, cudaMemcpyHostToDevice, ctx->stream));
    return CALZ_OK;
}
int stage_out(calz_ctx* ctx, const Staged& s, int64_t n, int cols, double* h, int64_t ldh) {
    CALZ_CUDA(ctx, cudaMemcpy2DAsync(h, (size_t)ldh * sizeof(double), s.d, (size_t)s.ld * sizeof(double),
                                     (size_t)n * sizeof(double), (size_t)cols, cudaMemcpyDeviceToHost, ctx->stream));
    CALZ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return CALZ_OK;
}
// all Q blocks of a cell array side by side in one staging buffer
int stage_blocks(calz_ctx* ctx, DevBuf& buf, int64_t n, int nblk, const double* const* Qblk, const int64_t* ldQ,
                 const int* mcols, std::vector<const double*>& dptr, std::vector<int64_t>& dld) {
    int tot = 0;
    for (int i = 0; i < nblk; ++i)
        if (!empty_block(Qblk, mcols, i)) tot += mcols[i];
    const int64_t ld = round_up(n, 32);
    CALZ_TRY(reserve(ctx, buf, (size_t)ld * std::max(tot, 1) * sizeof(double)));
    dptr.assign(nblk, nullptr);
    dld.assign(nblk, ld);
    int col = 0;
    for (int i = 0; i < nblk; ++i) {
        if (empty_block(Qblk, mcols, i)) continue;
        double* d = (double*)buf.p + (size_t)col * ld;
        CALZ_CUDA(ctx, cudaMemcpy2DAsync(d, (size_t)ld * sizeof(double), Qblk[i], (size_t)ldQ[i] * sizeof(double),
                                         (size_t)n * sizeof(double), (size_t)mcols[i], cudaMemcpyHostToDevice, ctx->stream));
        dptr[i] = d;
        col += mcols[i];
    }
    return CALZ_OK;
}
}  // namespace

int calz_tsqr_host(calz_ctx* ctx, int64_t n, int c, const double* A, int64_t ldA, double* Q, int64_t ldQ, double* R) {
    if (!ctx || !A || !Q || !R || n < 1 || c < 1) return set_error(ctx, CALZ_ERR_BADARG, "calz_tsqr_host: bad arguments");
    CALZ_CUDA(ctx, cudaSetDevice(ctx->device));
    Staged a, q;
    CALZ_TRY(stage_in(ctx, ctx->work[1], n, c, A, ldA, &a, true));
    CALZ_TRY(stage_in(ctx, ctx->work[2], n, c, nullptr, 0, &q, false));
    CALZ_TRY(calz_tsqr(ctx, n, c, a.d, a.ld, q.d, q.ld, R));
    return stage_out(ctx, q, n, c, Q, ldQ);
}

int calz_cholqr_host(calz_ctx* ctx, int64_t n, int c, const double* X, int64_t ldX, double* Q, int64_t ldQ, double* R, int* info) {
    if (!ctx || !X || !Q || !R || n < 1 || c < 1) return set_error(ctx, CALZ_ERR_BADARG, "calz_cholqr_host: bad arguments");
    CALZ_CUDA(ctx, cudaSetDevice(ctx->device));
    Staged a, q;
    CALZ_TRY(stage_in(ctx, ctx->work[1], n, c, X, ldX, &a, true));
    CALZ_TRY(stage_in(ctx, ctx->work[2], n, c, nullptr, 0, &q, false));
    CALZ_TRY(calz_cholqr(ctx, n, c, a.d, a.ld, q.d, q.ld, R, info));
    return stage_out(ctx, q, n, c, Q, ldQ);
}

int calz_normalize_host(calz_ctx* ctx, int64_t n, int c, const double* X, int64_t ldX, int backend, double tol, double* Q,
                        int64_t ldQ, double* R, int* rank) {
    if (!ctx || !X || !Q || !R || n < 1 || c < 1) return set_error(ctx, CALZ_ERR_BADARG, "calz_normalize_host: bad arguments");
    CALZ_CUDA(ctx, cudaSetDevice(ctx->device));
    Staged a, q;
    CALZ_TRY(stage_in(ctx, ctx->work[1], n, c, X, ldX, &a, true));
    CALZ_TRY(stage_in(ctx, ctx->work[2], n, c, nullptr, 0, &q, false));
    CALZ_TRY(calz_normalize(ctx, n, c, a.d, a.ld, backend, tol, q.d, q.ld, R, rank));
    return stage_out(ctx, q, n, c, Q, ldQ);
}

int calz_project_host(calz_ctx* ctx, int64_t n, int nblk, const double* const* Qblk, const int64_t* ldQ, const int* mcols,
                      int c, double* X, int64_t ldX, int doreorth, double* const* Rblk) {
    if (!ctx || !X || n < 1 || c < 1) return set_error(ctx, CALZ_ERR_BADARG, "calz_project_host: bad arguments");
    CALZ_CUDA(ctx, cudaSetDevice(ctx->device));
    Staged x;
    std::vector<const double*> dq;
    std::vector<int64_t> dld;
    CALZ_TRY(stage_in(ctx, ctx->work[1], n, c, X, ldX, &x, true));
    CALZ_TRY(stage_blocks(ctx, ctx->work[3], n, nblk, Qblk, ldQ, mcols, dq, dld));
    CALZ_TRY(calz_project(ctx, n, nblk, dq.data(), dld.data(), mcols, c, x.d, x.ld, doreorth, Rblk));
    return stage_out(ctx, x, n, c, X, ldX);
}

int calz_project_and_normalize_host(calz_ctx* ctx, int64_t n, int nblk, const double* const* Qblk, const int64_t* ldQ,
                                    const int* mcols, int c, const double* X, int64_t ldX, int doreorth, int backend,
                                    double* QZ, int64_t ldQZ, double* const* Rblk, double* Rlast, int* second_pass, int* rank) {
    if (!ctx || !X || !QZ || n < 1 || c < 1) return set_error(ctx, CALZ_ERR_BADARG, "calz_project_and_normalize_host: bad arguments");
    CALZ_CUDA(ctx, cudaSetDevice(ctx->device));
    Staged x, q;
    std::vector<const double*> dq;
    std::vector<int64_t> dld;
    CALZ_TRY(stage_in(ctx, ctx->work[1], n, c, X, ldX, &x, true));
    CALZ_TRY(stage_in(ctx, ctx->work[2], n, c, nullptr, 0, &q, false));
    CALZ_TRY(stage_blocks(ctx, ctx->work[3], n, nblk, Qblk, ldQ, mcols, dq, dld));
    int st = calz_project_and_normalize(ctx, n, nblk, dq.data(), dld.data(), mcols, c, x.d, x.ld, doreorth, backend, q.d, q.ld,
                                        Rblk, Rlast, second_pass, rank);
    if (st != CALZ_OK && st != CALZ_ERR_CHOL) return st;
    int st2 = stage_out(ctx, q, n, c, QZ, ldQZ);
    return st != CALZ_OK ? st : st2;
}

}  // extern "C"

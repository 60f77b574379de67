// Tall-skinny fp64 building blocks shared by the orthogonalisation entry points (tsops.cu, tsqr.cu).
#pragma once
#include "calz_internal.h"

namespace calz {

constexpr int kMaxC = 32;          // widest block being orthogonalised (s <= 31)
constexpr int kTsThreads = 256;

// A "panel list": virtual n x M matrix made of up to 4 column panels (a MATLAB cell array row, or [Q X]).
struct Panels {
    const double* ptr[4];
    long long ld[4];
    int ncols[4];
    int count;
    int total;
};
inline Panels one_panel(const double* p, int64_t ld, int ncols) {
    Panels P{};
    P.ptr[0] = p; P.ld[0] = ld; P.ncols[0] = ncols; P.count = 1; P.total = ncols;
    return P;
}
inline void add_panel(Panels& P, const double* p, int64_t ld, int ncols) {
    P.ptr[P.count] = p; P.ld[P.count] = ld; P.ncols[P.count] = ncols; P.count++; P.total += ncols;
}

// C = A'B (A = panel list n x M, B n x c) -> C_dev (M x c, leading dimension ldC), summed over ranks.
// `pred`/`want`: if pred != NULL the kernels run only when *pred == want (device-side control flow).
int tsmm_tn(calz_ctx* ctx, int64_t n, const Panels& A, const double* B, int64_t ldB, int c, double* C_dev,
            int ldC, bool same, const int* pred, int want, bool allreduce);

// Y = X - Q*C   (Q n x M single panel, C device M x c with ld ldC).  Y may alias X.
int ts_update(calz_ctx* ctx, int64_t n, const double* Q, int64_t ldQ, int M, const double* C_dev, int ldC,
              const double* X, int64_t ldX, int c, double* Y, int64_t ldY, const int* pred, int want);

// Q = X / R  (R device c x c upper, ld c).  Q may alias X.
int ts_trsolve(calz_ctx* ctx, int64_t n, int c, const double* X, int64_t ldX, const double* R, double* Q,
               int64_t ldQ, const int* pred, int want);

// Rfin = *sel ? R_b : R_a;  *reorth_out = *sel ? *cond_b : *cond_a  (cond_* from chol_small, may be NULL)
int select_r(calz_ctx* ctx, int c, const double* R_a, const double* R_b, const int* sel, const int* cond_a,
             const int* cond_b, double* Rfin, int* reorth_out);
// Rfin = Rb * Rfin (upper triangular product), predicated
int rmul_upper(calz_ctx* ctx, int c, const double* Rb, double* Rfin, const int* pred, int want);

// R = chol(G) on the device (one warp), optional norm-drop decision of projectAndNormalize.m:45-52:
// *flag_out = max_i |sqrt(nb2[i*nb2_stride]) - ||R(:,i)|| | / sqrt(nb2[...]) > 0.5.  info_out: 0 or failing pivot.
// adaptive (CALZ_QR_CHOLQR2): shifted retry on breakdown (info_out = -#shifts), cond_out = "needs another CholQR pass".
int chol_small(calz_ctx* ctx, int c, const double* G_dev, double* R_dev, int* info_out, const double* nb2,
               int nb2_stride, int* flag_out, const int* pred, int want, bool adaptive = false, int* cond_out = nullptr,
               int ldG = 0);

// fused small-matrix step of the tile pipeline: R1 = chol(G_Y), norm-drop test, R2 = chol(G_Y - C2'C2) if it fires, Rf, flags
int chol_pan(calz_ctx* ctx, int c, int M, const double* S2, int ldS, const double* nb2, int nb2_stride, bool adaptive, double* R1,
             double* R2, double* Rf, int* flags);
// final fused pass: QZ = ((X - Q*C1) - [*flag2] Q*C2) / R in one sweep over [Q | X] (tiles.cu)
int tile_update_solve(calz_ctx* ctx, int64_t n, const double* Q, int64_t ldQ, int M, const double* X, int64_t ldX, int c,
                      const double* C1_dev, int ldC1, const double* C2_dev, int ldC2, const int* flag2, const double* R_dev, double* QZ,
                      int64_t ldQZ);

// fused TMA-tile passes of projectAndNormalize (tiles.cu)
bool tile_path_ok(int64_t n, const double* Q, int64_t ldQ, int M, const double* X, int64_t ldX, int c, const double* Y, int64_t ldY);
int tile_panel_width(int c);        // widest Q panel per pass (0: the block is too wide for the tile kernels)
// S_dev == NULL (update modes): a pure update, nothing contracted.  Panels wider than 16 columns: modes 0 and 2 only.
int tile_pass(calz_ctx* ctx, int mode, int64_t n, const double* Q, int64_t ldQ, int M, const double* X, int64_t ldX, int c,
              const double* C_dev, int ldC, double* Y, int64_t ldY, double* S_dev, int ldS, const int* pred, int want, bool allreduce);

int norm_drop_from_gram(calz_ctx* ctx, int c, const double* G_dev, int ldG, const double* nb2, int nb2_stride, int* flag_out);
// same decision for a backend that already has R on the device (TSQR)
int norm_drop_decision(calz_ctx* ctx, int c, const double* R_dev, const double* nb2, int nb2_stride, int* flag_out);

// TSQR (tsqr.cu): R factor only / apply.  R_dev c x c upper with diag >= 0 (tsqr.m:9-11).
int tsqr_factor(calz_ctx* ctx, int64_t n, int c, const double* A, int64_t ldA, double* R_dev, const int* pred, int want);
int tsqr_form_q(calz_ctx* ctx, int64_t n, int c, const double* A, int64_t ldA, double* Q, int64_t ldQ);

}  // namespace calz

/*
 * calz.h -- C ABI of the B200-native CA-Lanczos hot path (libcalz.so).
 *
 * The reference (magnusgrandin/ca-lanczos) is pure MATLAB: its drivers call the block kernels by plain
 * function name (ca_lanczos.m:113,116,178,187,197; restarted_ca_lanczos.m:274,277,313,315,324,330,333).
 * A drop-in is therefore a same-named MEX gateway (mex/ *.cpp) that forwards to the entry points below;
 * INTEGRATION.md shows the gateway for each.  Every entry point cites the reference interface it replaces.
 *
 * Conventions
 *   - plain C: pointers + sizes, no C++ types, no exceptions cross this boundary;
 *   - every function returns an int status (CALZ_OK == 0); calz_last_error() gives the text;
 *   - dense data is IEEE fp64, column-major, explicit leading dimension (MATLAB layout);
 *   - "_host" entry points take HOST pointers and are synchronous (what the MEX gateways call);
 *     the un-suffixed compute entry points take DEVICE pointers and run on the context's stream;
 *   - small results (R factors, flags) are always written to HOST pointers;
 *   - there is no CPU fallback: without a CUDA device calz_init fails with CALZ_ERR_CUDA.
 */
#ifndef CALZ_H
#define CALZ_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif
#if defined(__GNUC__)
#pragma GCC visibility push(default)   /* the library is built with -fvisibility=hidden: only calz_* is exported */
#endif

#define CALZ_OK                 0
#define CALZ_ERR_BADARG         1
#define CALZ_ERR_CUDA           2
#define CALZ_ERR_NCCL           3
#define CALZ_ERR_CHOL           4   /* Gram matrix not positive definite (MATLAB chol error, cholqr.m:6) */
#define CALZ_ERR_ALLOC          5
#define CALZ_ERR_UNSUPPORTED    6
#define CALZ_ERR_CLOSURE        7   /* supplied rows do not cover the level-(s-1) closure of the owned rows */
#define CALZ_ERR_SHIFT          8   /* matrix_powers_newton.m:36-39: negative imaginary shift at k==1 */

/* sparse device layouts (north_star: CSR plus SELL-C-sigma) */
#define CALZ_LAYOUT_AUTO        0
#define CALZ_LAYOUT_CSR         1
#define CALZ_LAYOUT_SELL        2
#define CALZ_LAYOUT_SELL_DICT   3   /* SELL-32 with one dictionary code byte per non-zero: needs <= 255 distinct
                                       (column offset, value) pairs (constant-coefficient stencils); lossless */

/* QR backend at the normalize.m:14 seam */
#define CALZ_QR_TSQR            0   /* tsqr.m:7-12   (reference default) */
#define CALZ_QR_CHOLQR          1   /* cholqr.m:3-8, single pass (orthogonality ~ kappa^2 eps) */
#define CALZ_QR_CHOLQR2         2   /* cholqr.m + an automatic second CholQR pass (R = R2*R1) when the on-device
                                       conditioning estimate min_j R_jj/||x_j|| says one pass is not
                                       Householder-accurate; identical kernels and results to CHOLQR otherwise */

typedef struct calz_ctx calz_ctx;   /* one per process / per GPU: device, stream, scratch, communicator */
typedef struct calz_mat calz_mat;   /* device sparse matrix: owned rows + level-s ghost closure */

/* ------------------------------------------------------------------ context ------------------------ */
int  calz_version(void);
int  calz_init(int device, calz_ctx** ctx);
int  calz_finalize(calz_ctx* ctx);
const char* calz_last_error(const calz_ctx* ctx);        /* ctx may be NULL: last global error */
int  calz_set_stream(calz_ctx* ctx, void* cuda_stream);  /* adopt a caller stream (NULL: own stream) */
void* calz_get_stream(calz_ctx* ctx);
int  calz_sync(calz_ctx* ctx);
/* number of kernels launched by this library since the last reset (bench.py "gpu_launches") */
int64_t calz_launch_count(calz_ctx* ctx, int reset);
/* knobs (INTEGRATION.md has the table with defaults): "mpk_patterns", "mpk_prefetch", "mpk_pair_phase", "mpk_halo_level",
 * "mpk_dict_mode", "mpk_persist", "mpk_tma_x", "mpk_l2_chunk_bytes", "mpk_fused_steps", "sell_dict", "sell_sigma", "csr_lanes",
 * "grid_mult", "tile_pipeline", "tile_panels", "pan_fused_solve", "fused_allreduce", "p2p",
 * "cholqr2_inv_thresh" (second pass when min_j R_jj/||x_j|| < 1/value; default 32).  Unknown key: CALZ_ERR_BADARG. */
int  calz_set_option(calz_ctx* ctx, const char* key, int64_t value);

/* ------------------------------------------------------------------ process-wide state (MEX gateways) -- */
/* The eight MEX gateways are eight shared objects in one MATLAB/Octave process.  They must share ONE context (one CUDA
 * context's worth of scratch) and ONE device-matrix cache, so both live inside libcalz.so, which the process maps once:
 * calz_shared_context creates the context on first use (device from CALZ_DEVICE, default 0) and returns the same one afterwards;
 * calz_shared_release destroys the cache and the context (mexAtExit). */
int  calz_shared_context(calz_ctx** ctx);
int  calz_shared_release(void);
/* Device copy of a MATLAB sparse matrix (CSC, 64-bit mwIndex: mxGetJc/mxGetIr/mxGetPr), uploaded once and reused across
 * calls (SpMV.m:3-5 "other data structures").  Keyed on (pr, n, nnz) AND a content fingerprint -- a hash of 4096 strided
 * samples of pr/ir/jc plus their ends -- because MATLAB reuses freed data pointers: `A = sparse(diag(a))` rebuilt in a loop
 * with constant n (test_restart_diagonal_matrices.m:23) would otherwise hit the previous matrix.  A stale entry is replaced.
 * calz_mat_cache_clear drops every cached matrix (explicit invalidation). */
int  calz_mat_cache_get_csc64(calz_ctx* ctx, int64_t n, const uint64_t* jc, const uint64_t* ir, const double* pr, int s_max,
                              int layout, calz_mat** mat);
int  calz_mat_cache_clear(calz_ctx* ctx);

/* ------------------------------------------------------------------ device blocks (handle mode) ------- */
/* An n x cols fp64 block that lives on the device (column-major, leading dimension ld = n rounded up to 32): what a gateway
 * hands back instead of a host array in handle mode (mex/README.md), so that V, Q and QZ never cross PCIe between calls.
 * The device pointer feeds the un-suffixed entry points (calz_mpk_newton, calz_project_and_normalize, ...). */
typedef struct calz_vec calz_vec;
int  calz_vec_create(calz_ctx* ctx, int64_t n, int cols, calz_vec** v);
int  calz_vec_destroy(calz_vec* v);
int  calz_vec_info(const calz_vec* v, double** dev, int64_t* n, int* cols, int64_t* ld);
int  calz_vec_upload(calz_vec* v, int col0, int cols, const double* host, int64_t ldh);        /* host -> columns [col0, col0+cols) */
int  calz_vec_download(const calz_vec* v, int col0, int cols, double* host, int64_t ldh);      /* synchronous */
/* dst(:, dcol0 : dcol0+cols) = src(:, scol0 : scol0+cols) on the device (MATLAB `Q(:,a:b) = Q_` on handles) */
int  calz_vec_copy(calz_vec* dst, int dcol0, const calz_vec* src, int scol0, int cols);

/* ------------------------------------------------------------------ multi-GPU plumbing --------------- */
/* One process per GPU.  The 128-byte id is created on rank 0 and shipped by the host framework
 * (torch.distributed broadcast); nccl_lib may be NULL (uses the libnccl.so.2 already in the process). */
int  calz_comm_unique_id(char id_out[128], const char* nccl_lib);
int  calz_comm_init(calz_ctx* ctx, int nranks, int rank, const char id[128], const char* nccl_lib);
int  calz_comm_rank(const calz_ctx* ctx, int* rank, int* nranks);

/* ------------------------------------------------------------------ partition / ghost plan (host only) */
/* Row partition: rank p owns [floor(p*n/P), floor((p+1)*n/P)).  bounds has P+1 entries. */
int  calz_partition_bounds(int64_t n, int P, int64_t* bounds);
/* Level sets of the pattern graph from owned rows [lo,hi): level_out[j] (length n_glob) = smallest k<=s
 * with j in R_k, else -1.  rowptr/colind describe rows [row_begin,row_end) of the GLOBAL matrix
 * (rowptr[0]==0, global column indices); they must cover R_{s-1}.  Pure host code, no GPU needed. */
int  calz_level_sets(int64_t n_glob, int64_t row_begin, int64_t row_end, const int64_t* rowptr,
                     const int32_t* colind, int64_t lo, int64_t hi, int s, int32_t* level_out);

/* ------------------------------------------------------------------ sparse matrix -------------------- */
/* Replaces the first argument of SpMV.m:6 ("other data structures", SpMV.m:3-5).
 * Rows [row_begin,row_end) of the n_glob x n_glob matrix in CSR (int64 rowptr starting at 0, int32 GLOBAL
 * column indices ascending within a row, fp64 values).  With a communicator of P ranks the context's rank
 * owns calz_partition_bounds rows and keeps the level-s_max ghost closure (PA1); P==1: the whole matrix
 * (row_begin=0,row_end=n_glob). */
int  calz_mat_create_csr(calz_ctx* ctx, int64_t n_glob, int64_t row_begin, int64_t row_end,
                         const int64_t* rowptr, const int32_t* colind, const double* val,
                         int s_max, int layout, calz_mat** mat);
/* MATLAB sparse (mxGetJc/mxGetIr/mxGetPr: CSC, 64-bit mwIndex).  Transposed on the host once, so A need not
 * be symmetric.  Single-GPU convenience for the MEX gateways. */
int  calz_mat_create_csc64(calz_ctx* ctx, int64_t n, const uint64_t* jc, const uint64_t* ir,
                           const double* pr, int s_max, int layout, calz_mat** mat);
int  calz_mat_destroy(calz_mat* mat);
/* what[] keys: "n_glob","n_own","n_loc","own_off","row_lo","row_hi","nnz_loc","layout","sell_padded_nnz",
 * "n_ghost","bandwidth" */
int  calz_mat_info(const calz_mat* mat, const char* what, int64_t* value);
/* ghost_s(p) (sorted global indices) and the per-peer receive lists; idx_out may be NULL to query count */
int  calz_mat_ghost_indices(const calz_mat* mat, int64_t* idx_out, int64_t* count);
int  calz_mat_recv_list(const calz_mat* mat, int peer, int64_t* idx_out, int64_t* count);
int  calz_mat_send_list(const calz_mat* mat, int peer, int64_t* idx_out, int64_t* count);

/* ------------------------------------------------------------------ matrix powers kernel ------------- */
/* SpMV.m:6-8            y = A*x                     (owned rows; x,y device, length n_own) */
int  calz_spmv(calz_mat* mat, const double* x, double* y);
/* matrix_powers_monomial.m:6-12   V(:,1)=A*q, V(:,i)=A*V(:,i-1)   -> V is n_own x s (q NOT included) */
int  calz_mpk_monomial(calz_mat* mat, const double* q, int s, double* V, int64_t ldV);
/* matrix_powers_newton.m:15-54    V(:,1)=v, V(:,k+1)=A*V(:,k)-re(l_k)*V(:,k) [+im(l_k)^2*V(:,k-1)]
 * -> V is n_own x (s+1).  shift_im may be NULL (real shifts).  modifiedp as in the reference (callers
 * pass 1, ca_lanczos.m:116); modifiedp==0 with complex shifts needs complex vectors: CALZ_ERR_UNSUPPORTED. */
int  calz_mpk_newton(calz_mat* mat, const double* v, int s, const double* shift_re, const double* shift_im,
                     int modifiedp, double* V, int64_t ldV);
/* Zero-copy variant: the basis stays in the matrix' own workspace (n_loc x (s+1), ghosts included);
 * *V points at the owned rows of column 0, *ldV is the workspace leading dimension.  monomial!=0 computes
 * the monomial basis (column 0 = q, so V(:,2:s+1) is matrix_powers_monomial's output). */
int  calz_mpk_inplace(calz_mat* mat, const double* v, int s, const double* shift_re, const double* shift_im,
                      int modifiedp, int monomial, double** V, int64_t* ldV);
/* host-pointer flavours (synchronous; H2D of the vector, D2H of the basis inside) */
int  calz_spmv_host(calz_mat* mat, const double* x, double* y);
int  calz_mpk_monomial_host(calz_mat* mat, const double* q, int s, double* V, int64_t ldV);
int  calz_mpk_newton_host(calz_mat* mat, const double* v, int s, const double* shift_re,
                          const double* shift_im, int modifiedp, double* V, int64_t ldV);

/* ------------------------------------------------------------------ block orthogonalisation ---------- */
/* All take the LOCAL row count n (owned rows of this rank); with a communicator the small Gram / R /
 * coefficient matrices are all-reduced, so every rank returns identical R factors.
 *
 * tsqr.m:7-12      [Q,R]=qr(A,0) with diag(R)>=0.  A n x c -> Q n x c (may alias A), R c x c (host, ld c) */
int  calz_tsqr(calz_ctx* ctx, int64_t n, int c, const double* A, int64_t ldA, double* Q, int64_t ldQ, double* R);
/* cholqr.m:3-8     G=X'X; R=chol(G); Q=X/R.  info: 0, or j>0 if the j-th pivot failed (CALZ_ERR_CHOL) */
int  calz_cholqr(calz_ctx* ctx, int64_t n, int c, const double* X, int64_t ldX, double* Q, int64_t ldQ,
                 double* R, int* info);
/* normalize.m:3-36 QR by `backend`, svd(R), rank = #{sigma_i > tol*sigma_1} (tol default 1e-8: pass <=0) */
int  calz_normalize(calz_ctx* ctx, int64_t n, int c, const double* X, int64_t ldX, int backend, double tol,
                    double* Q, int64_t ldQ, double* R, int* rank);
/* project.m:7-58   for each block i (in order; mcols[i]==0 or Qblk[i]==NULL is an empty cell):
 *                  R{i}=Q{i}'*X; X=X-Q{i}*R{i}.  X is updated in place; Rblk[i] host, mcols[i] x c.
 *                  doreorth!=0 restates project.m:40-57 (inverted criterion kept). */
int  calz_project(calz_ctx* ctx, int64_t n, int nblk, const double* const* Qblk, const int64_t* ldQ,
                  const int* mcols, int c, double* X, int64_t ldX, int doreorth, double* const* Rblk);
/* projectAndNormalize.m:3-90   QZ n x c (must not alias X), Rblk[i] host mcols[i] x c (pass-1 + pass-2
 * coefficients, :71-73), Rlast host c x c (R of the LAST normalize), second_pass (replaces disp('second')
 * :62), rank (normalize.m:18-24). */
int  calz_project_and_normalize(calz_ctx* ctx, int64_t n, int nblk, const double* const* Qblk,
                                const int64_t* ldQ, const int* mcols, int c, const double* X, int64_t ldX,
                                int doreorth, int backend, double* QZ, int64_t ldQZ, double* const* Rblk,
                                double* Rlast, int* second_pass, int* rank);
/* Asynchronous variant for the device-resident pipeline: enqueues the whole call and returns a ticket at once (up
 * to 8 calls in flight); calz_pan_collect waits for that call only and unpacks its small results.  The rare CholQR2
 * refinement cannot run behind the caller's back: *needs_refine=1 tells the caller to redo that block synchronously. */
int  calz_project_and_normalize_async(calz_ctx* ctx, int64_t n, int nblk, const double* const* Qblk,
                                      const int64_t* ldQ, const int* mcols, int c, const double* X, int64_t ldX,
                                      int doreorth, int backend, double* QZ, int64_t ldQZ, int* ticket);
int  calz_pan_collect(calz_ctx* ctx, int ticket, double* const* Rblk, double* Rlast, int* second_pass, int* rank,
                      int* needs_refine);
/* host-pointer flavours (synchronous) */
int  calz_tsqr_host(calz_ctx* ctx, int64_t n, int c, const double* A, int64_t ldA, double* Q, int64_t ldQ, double* R);
int  calz_cholqr_host(calz_ctx* ctx, int64_t n, int c, const double* X, int64_t ldX, double* Q, int64_t ldQ,
                      double* R, int* info);
int  calz_normalize_host(calz_ctx* ctx, int64_t n, int c, const double* X, int64_t ldX, int backend, double tol,
                         double* Q, int64_t ldQ, double* R, int* rank);
int  calz_project_host(calz_ctx* ctx, int64_t n, int nblk, const double* const* Qblk, const int64_t* ldQ,
                       const int* mcols, int c, double* X, int64_t ldX, int doreorth, double* const* Rblk);
int  calz_project_and_normalize_host(calz_ctx* ctx, int64_t n, int nblk, const double* const* Qblk,
                                     const int64_t* ldQ, const int* mcols, int c, const double* X, int64_t ldX,
                                     int doreorth, int backend, double* QZ, int64_t ldQZ, double* const* Rblk,
                                     double* Rlast, int* second_pass, int* rank);

/* ------------------------------------------------------------------ building blocks (device) --------- */
/* C = A'B  (A n x m, B n x c -> C m x c, DEVICE, ld m), all-reduced over the communicator: cholqr.m:5,
 * project.m:34.  The fp64 DMMA tall-skinny contraction, exposed for the Gram roofline measurement. */
int  calz_gram(calz_ctx* ctx, int64_t n, int m, const double* A, int64_t ldA, int c, const double* B,
               int64_t ldB, double* C_dev);
/* Y = X - Q*C  (Q n x m, C m x c on the HOST, X n x c or NULL for zero; Y may alias X).  The tall-skinny update of
 * project.m:35, exposed for the callers' own O(n) work: the three-term recurrence of lanczos.m:103-110 and the Ritz
 * vector assembly Q*Vp of restarted_ca_lanczos.m:135-139 / ca_lanczos.m:94 (pass C = -Vp, X = NULL). */
int  calz_block_axpy(calz_ctx* ctx, int64_t n, int m, const double* Q, int64_t ldQ, int c, const double* C_host,
                     const double* X, int64_t ldX, double* Y, int64_t ldY);
/* Orthogonality diagnostics of the drivers on the device; blocks as in calz_project (device pointers, the
 * virtual matrix Q = [Q_1 ... Q_nblk] has `tot` columns).  One pass of DMMA Gram products, all-reduced; the tot x tot
 * arithmetic is done on the host.
 *   CALZ_ORTH_FRO        norm(eye(tot) - Q'*Q, 'fro')                          restarted_ca_lanczos.m:165-168
 *   CALZ_ORTH_LASTBLOCK  compute_orth_err(Q, s) of ca_lanczos.m:99-107: tot > s+1: max|Q(:,1:tot-s-1)'*Q(:,tot-s:tot)|,
 *                        else max|Q'*Q - eye(s+1)|  */
#define CALZ_ORTH_FRO        0
#define CALZ_ORTH_LASTBLOCK  1
int  calz_orth_error(calz_ctx* ctx, int64_t n, int nblk, const double* const* Qblk, const int64_t* ldQ, const int* mcols,
                     int mode, int s, double* err);
/* measured peak of the fp64 tensor pipe (register-resident mma.sync.m8n8k4.f64 chains), TFLOP/s: the denominator of
 * the Gram kernels' "fraction of the fp64 tensor pipe" */
int  calz_dmma_peak(calz_ctx* ctx, double* tflops);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#ifdef __cplusplus
}
#endif
#endif /* CALZ_H */

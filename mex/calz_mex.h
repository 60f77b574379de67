// Shared plumbing of the MEX gateways: the process-wide libcalz context and device-matrix cache (both owned by libcalz.so),
// status -> mexErrMsgIdAndTxt mapping, QR backend selection, and the calz_vec handle mode.
//
// Each gateway is a same-named drop-in for a reference .m file (matrix_powers_monomial.m, ... ); MATLAB/Octave
// resolve the call by name, MEX before .m on the same path, so the reference drivers (ca_lanczos.m,
// restarted_ca_lanczos.m) pick them up unchanged.  No numerics live here -- everything goes through the C ABI.
#pragma once
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "calz.h"
#include "mex.h"

// ---- process-wide state lives INSIDE libcalz.so (calz_shared_context / calz_mat_cache_get_csc64): the eight gateways are eight
//      shared objects, a header-local static here would give each of them its own CUDA context and its own matrix cache.
inline void calz_mex_cleanup(void) { calz_shared_release(); }

inline void calz_mex_fail(int status, const char* where) {
    if (status == CALZ_OK) return;
    static const char* ids[] = {"calanczos:ok", "calanczos:badarg", "calanczos:cuda", "calanczos:nccl", "calanczos:chol",
                                "calanczos:alloc", "calanczos:unsupported", "calanczos:closure", "calanczos:shift"};
    const char* id = (status >= 0 && status <= 8) ? ids[status] : "calanczos:unknown";
    calz_ctx* ctx = nullptr;
    calz_shared_context(&ctx);
    // mexErrMsgIdAndTxt longjmps: nothing owned by C++ may be live past this point in the caller
    mexErrMsgIdAndTxt(id, "%s: %s", where, calz_last_error(ctx));
}

inline calz_ctx* calz_mex_context() {
    static bool registered = false;           // per gateway: mexAtExit / mexLock are per MEX file; the release is idempotent
    calz_ctx* ctx = nullptr;
    int st = calz_shared_context(&ctx);
    if (st != CALZ_OK) mexErrMsgIdAndTxt("calanczos:cuda", "calz_init: %s", calz_last_error(nullptr));     // no CPU fallback
    if (!registered) {
        mexAtExit(calz_mex_cleanup);
        mexLock();
        registered = true;
    }
    return ctx;
}

// normalize.m:14 seam: tsqr (reference default) unless CALZ_QR_BACKEND=cholqr
inline int calz_mex_backend() {
    const char* b = getenv("CALZ_QR_BACKEND");
    if (b && !strcmp(b, "cholqr")) return CALZ_QR_CHOLQR;
    if (b && !strcmp(b, "cholqr2")) return CALZ_QR_CHOLQR2;
    return CALZ_QR_TSQR;
}

// Device copy of a MATLAB sparse matrix, uploaded once and reused across calls (SURVEY.md §7 "hard parts"); the cache is keyed
// on the data pointer, n, nnz AND a content fingerprint (MATLAB reuses freed addresses: test_restart_diagonal_matrices.m:23).
inline calz_mat* calz_mex_matrix(const mxArray* A) {
    if (!mxIsSparse(A) || !mxIsDouble(A) || mxIsComplex(A) || mxGetM(A) != mxGetN(A))
        mexErrMsgIdAndTxt("calanczos:badarg", "A must be a real sparse square matrix");
    calz_ctx* ctx = calz_mex_context();
    static_assert(sizeof(mwIndex) == 8, "64-bit mwIndex expected (-largeArrayDims)");
    calz_mat* m = nullptr;
    int st = calz_mat_cache_get_csc64(ctx, (int64_t)mxGetN(A), (const uint64_t*)mxGetJc(A), (const uint64_t*)mxGetIr(A), mxGetPr(A),
                                      /*s_max=*/32, CALZ_LAYOUT_AUTO, &m);
    calz_mex_fail(st, "calz_mat_cache_get_csc64");
    return m;
}

// ---- handle mode (mex/README.md): a `calz_vec` MATLAB object (mex/calz_vec.m) stands for an n x cols block that lives on the
//      device; its properties are the libcalz handle, the row count and the column window [col0, col0+cols) of a view.
struct CalzMexVec {
    calz_vec* h = nullptr;
    double* dev = nullptr;         // device pointer of the view's first column
    int64_t n = 0, ld = 0;
    int cols = 0, col0 = 0;        // the view: columns [col0, col0 + cols) of the block behind h
};
inline bool calz_mex_is_vec(const mxArray* a) { return a && mxIsClass(a, "calz_vec"); }
inline CalzMexVec calz_mex_vec(const mxArray* a) {
    CalzMexVec v;
    const mxArray* h = mxGetProperty(a, 0, "h");
    const mxArray* c0 = mxGetProperty(a, 0, "col0");
    const mxArray* nc = mxGetProperty(a, 0, "cols");
    if (!h || !c0 || !nc) mexErrMsgIdAndTxt("calanczos:badarg", "not a calz_vec object");
    v.h = (calz_vec*)(uintptr_t)(*(const uint64_t*)mxGetData(h));
    double* base = nullptr;
    int all = 0;
    if (calz_vec_info(v.h, &base, &v.n, &all, &v.ld) != CALZ_OK) mexErrMsgIdAndTxt("calanczos:badarg", "stale calz_vec handle");
    const int col0 = (int)mxGetScalar(c0);
    v.cols = (int)mxGetScalar(nc);
    if (col0 < 0 || v.cols < 1 || col0 + v.cols > all) mexErrMsgIdAndTxt("calanczos:badarg", "calz_vec view out of range");
    v.dev = base + (size_t)col0 * v.ld;
    v.col0 = col0;
    return v;
}
// a fresh device block wrapped into a calz_vec object (the MATLAB constructor takes the handle, n and the column count)
inline mxArray* calz_mex_new_vec(calz_ctx* ctx, size_t n, int cols, CalzMexVec* out) {
    calz_vec* h = nullptr;
    calz_mex_fail(calz_vec_create(ctx, (int64_t)n, cols, &h), "calz_vec_create");
    mxArray* args[3];
    args[0] = mxCreateNumericMatrix(1, 1, mxUINT64_CLASS, mxREAL);
    *(uint64_t*)mxGetData(args[0]) = (uint64_t)(uintptr_t)h;
    args[1] = mxCreateDoubleScalar((double)n);
    args[2] = mxCreateDoubleScalar((double)cols);
    mxArray* obj = nullptr;
    if (mexCallMATLAB(1, &obj, 3, args, "calz_vec") != 0) mexErrMsgIdAndTxt("calanczos:badarg", "calz_vec constructor failed");
    out->h = h; out->n = (int64_t)n; out->cols = cols;
    int all;
    calz_vec_info(h, &out->dev, nullptr, &all, &out->ld);
    return obj;
}

// MATLAB cell array of blocks -> pointer / leading-dimension / column-count arrays (empty cells stay empty)
struct CalzMexCell {
    std::vector<const double*> ptr;
    std::vector<int64_t> ld;
    std::vector<int> mcols;
    bool on_device = false;         // at least one block is a calz_vec: every non-empty block must be (checked by the caller)
};
inline void calz_mex_cell(const mxArray* Q, size_t n, CalzMexCell& out) {
    if (!mxIsCell(Q)) mexErrMsgIdAndTxt("calanczos:badarg", "Input Q (arg 1) must be cell (block) array.");   // project.m:12-15
    const size_t nb = mxGetNumberOfElements(Q);
    out.ptr.assign(nb, nullptr); out.ld.assign(nb, (int64_t)n); out.mcols.assign(nb, 0);
    for (size_t i = 0; i < nb; ++i) {
        const mxArray* Qi = mxGetCell(Q, i);
        if (calz_mex_is_vec(Qi)) {                                  // handle mode: the block already lives on the device
            CalzMexVec v = calz_mex_vec(Qi);
            if ((size_t)v.n != n) mexErrMsgIdAndTxt("calanczos:badarg", "Q{%d} has the wrong number of rows", (int)i + 1);
            out.ptr[i] = v.dev; out.ld[i] = v.ld; out.mcols[i] = v.cols; out.on_device = true;
            continue;
        }
        if (!Qi || mxIsEmpty(Qi)) continue;
        if (mxGetM(Qi) != n || !mxIsDouble(Qi) || mxIsComplex(Qi))
            mexErrMsgIdAndTxt("calanczos:badarg", "Q{%d} must be a real n-by-m double matrix", (int)i + 1);
        out.ptr[i] = mxGetPr(Qi); out.mcols[i] = (int)mxGetN(Qi);
    }
}

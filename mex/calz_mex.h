// Shared plumbing of the MEX gateways: one process-wide libcalz context, a device-matrix cache keyed on the
// MATLAB sparse array's data pointer, status -> mexErrMsgIdAndTxt mapping, QR backend selection.
//
// Each gateway is a same-named drop-in for a reference .m file (matrix_powers_monomial.m, ... ); MATLAB/Octave
// resolve the call by name, MEX before .m on the same path, so the reference drivers (ca_lanczos.m,
// restarted_ca_lanczos.m) pick them up unchanged.  No numerics live here -- everything goes through the C ABI.
#pragma once
#include <stdlib.h>
#include <string.h>

#include <map>
#include <tuple>
#include <vector>

#include "calz.h"
#include "mex.h"

struct CalzMexState {
    calz_ctx* ctx = nullptr;
    std::map<std::tuple<const void*, size_t, size_t>, calz_mat*> mats;
};
inline CalzMexState& calz_mex_state() {
    static CalzMexState s;
    return s;
}

inline void calz_mex_cleanup(void) {
    CalzMexState& s = calz_mex_state();
    for (auto& kv : s.mats) calz_mat_destroy(kv.second);
    s.mats.clear();
    if (s.ctx) calz_finalize(s.ctx);
    s.ctx = nullptr;
}

inline void calz_mex_fail(int status, const char* where) {
    if (status == CALZ_OK) return;
    static const char* ids[] = {"calanczos:ok", "calanczos:badarg", "calanczos:cuda", "calanczos:nccl", "calanczos:chol",
                                "calanczos:alloc", "calanczos:unsupported", "calanczos:closure", "calanczos:shift"};
    const char* id = (status >= 0 && status <= 8) ? ids[status] : "calanczos:unknown";
    // mexErrMsgIdAndTxt longjmps: nothing owned by C++ may be live past this point in the caller
    mexErrMsgIdAndTxt(id, "%s: %s", where, calz_last_error(calz_mex_state().ctx));
}

inline calz_ctx* calz_mex_context() {
    CalzMexState& s = calz_mex_state();
    if (!s.ctx) {
        const char* dev = getenv("CALZ_DEVICE");
        int st = calz_init(dev ? atoi(dev) : 0, &s.ctx);
        if (st != CALZ_OK) calz_mex_fail(st, "calz_init");     // no CPU fallback
        mexAtExit(calz_mex_cleanup);
        mexLock();
    }
    return s.ctx;
}

// normalize.m:14 seam: tsqr (reference default) unless CALZ_QR_BACKEND=cholqr
inline int calz_mex_backend() {
    const char* b = getenv("CALZ_QR_BACKEND");
    if (b && !strcmp(b, "cholqr")) return CALZ_QR_CHOLQR;
    if (b && !strcmp(b, "cholqr2")) return CALZ_QR_CHOLQR2;
    return CALZ_QR_TSQR;
}

// Device copy of a MATLAB sparse matrix, uploaded once and reused across calls (SURVEY.md §7 "hard parts").
inline calz_mat* calz_mex_matrix(const mxArray* A) {
    if (!mxIsSparse(A) || !mxIsDouble(A) || mxIsComplex(A) || mxGetM(A) != mxGetN(A))
        mexErrMsgIdAndTxt("calanczos:badarg", "A must be a real sparse square matrix");
    calz_ctx* ctx = calz_mex_context();
    CalzMexState& s = calz_mex_state();
    const size_t n = mxGetN(A), nnz = (size_t)mxGetJc(A)[n];
    auto key = std::make_tuple((const void*)mxGetPr(A), n, nnz);
    auto it = s.mats.find(key);
    if (it != s.mats.end()) return it->second;
    if (s.mats.size() >= 4) {                                   // small cache: drop everything when it fills up
        for (auto& kv : s.mats) calz_mat_destroy(kv.second);
        s.mats.clear();
    }
    calz_mat* m = nullptr;
    static_assert(sizeof(mwIndex) == 8, "64-bit mwIndex expected (-largeArrayDims)");
    int st = calz_mat_create_csc64(ctx, (int64_t)n, (const uint64_t*)mxGetJc(A), (const uint64_t*)mxGetIr(A), mxGetPr(A),
                                   /*s_max=*/32, CALZ_LAYOUT_AUTO, &m);
    calz_mex_fail(st, "calz_mat_create_csc64");
    s.mats[key] = m;
    return m;
}

// MATLAB cell array of blocks -> pointer / leading-dimension / column-count arrays (empty cells stay empty)
struct CalzMexCell {
    std::vector<const double*> ptr;
    std::vector<int64_t> ld;
    std::vector<int> mcols;
};
inline void calz_mex_cell(const mxArray* Q, size_t n, CalzMexCell& out) {
    if (!mxIsCell(Q)) mexErrMsgIdAndTxt("calanczos:badarg", "Input Q (arg 1) must be cell (block) array.");   // project.m:12-15
    const size_t nb = mxGetNumberOfElements(Q);
    out.ptr.assign(nb, nullptr); out.ld.assign(nb, (int64_t)n); out.mcols.assign(nb, 0);
    for (size_t i = 0; i < nb; ++i) {
        const mxArray* Qi = mxGetCell(Q, i);
        if (!Qi || mxIsEmpty(Qi)) continue;
        if (mxGetM(Qi) != n || !mxIsDouble(Qi) || mxIsComplex(Qi))
            mexErrMsgIdAndTxt("calanczos:badarg", "Q{%d} must be a real n-by-m double matrix", (int)i + 1);
        out.ptr[i] = mxGetPr(Qi); out.mcols[i] = (int)mxGetN(Qi);
    }
}

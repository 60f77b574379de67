// [Q,R] = tsqr(A)                                 drop-in for tsqr.m:7-12  (diag(R) >= 0)
// Handle mode: A a calz_vec => Q a calz_vec (R is always a host matrix).
#include "calz_mex.h"
void mexFunction(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]) {
    if (nrhs != 1 || nlhs > 2) mexErrMsgIdAndTxt("calanczos:badarg", "usage: [Q,R] = tsqr(A)");
    calz_ctx* ctx = calz_mex_context();
    mxArray *Q, *R;
    if (calz_mex_is_vec(prhs[0])) {
        CalzMexVec A = calz_mex_vec(prhs[0]), Qv;
        Q = calz_mex_new_vec(ctx, (size_t)A.n, A.cols, &Qv);
        R = mxCreateDoubleMatrix(A.cols, A.cols, mxREAL);
        calz_mex_fail(calz_tsqr(ctx, A.n, A.cols, A.dev, A.ld, Qv.dev, Qv.ld, mxGetPr(R)), "tsqr");
    } else {
        const size_t n = mxGetM(prhs[0]), c = mxGetN(prhs[0]);
        Q = mxCreateDoubleMatrix(n, c, mxREAL);
        R = mxCreateDoubleMatrix(c, c, mxREAL);
        calz_mex_fail(calz_tsqr_host(ctx, (int64_t)n, (int)c, mxGetPr(prhs[0]), (int64_t)n, mxGetPr(Q), (int64_t)n, mxGetPr(R)), "tsqr");
    }
    plhs[0] = Q;
    if (nlhs > 1) plhs[1] = R; else mxDestroyArray(R);
}

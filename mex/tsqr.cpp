// [Q,R] = tsqr(A)                                 drop-in for tsqr.m:7-12  (diag(R) >= 0)
#include "calz_mex.h"
void mexFunction(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]) {
    if (nrhs != 1 || nlhs > 2) mexErrMsgIdAndTxt("calanczos:badarg", "usage: [Q,R] = tsqr(A)");
    calz_ctx* ctx = calz_mex_context();
    const size_t n = mxGetM(prhs[0]), c = mxGetN(prhs[0]);
    mxArray* Q = mxCreateDoubleMatrix(n, c, mxREAL);
    mxArray* R = mxCreateDoubleMatrix(c, c, mxREAL);
    calz_mex_fail(calz_tsqr_host(ctx, (int64_t)n, (int)c, mxGetPr(prhs[0]), (int64_t)n, mxGetPr(Q), (int64_t)n, mxGetPr(R)), "tsqr");
    plhs[0] = Q;
    if (nlhs > 1) plhs[1] = R; else mxDestroyArray(R);
}

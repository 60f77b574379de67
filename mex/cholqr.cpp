// [Q,R] = cholqr(X)                               drop-in for cholqr.m:3-8  (errors like chol when X'X is not PD)
#include "calz_mex.h"
void mexFunction(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]) {
    if (nrhs != 1 || nlhs > 2) mexErrMsgIdAndTxt("calanczos:badarg", "usage: [Q,R] = cholqr(X)");
    calz_ctx* ctx = calz_mex_context();
    const size_t n = mxGetM(prhs[0]), c = mxGetN(prhs[0]);
    mxArray* Q = mxCreateDoubleMatrix(n, c, mxREAL);
    mxArray* R = mxCreateDoubleMatrix(c, c, mxREAL);
    int info = 0;
    calz_mex_fail(calz_cholqr_host(ctx, (int64_t)n, (int)c, mxGetPr(prhs[0]), (int64_t)n, mxGetPr(Q), (int64_t)n, mxGetPr(R), &info),
                  "cholqr");
    plhs[0] = Q;
    if (nlhs > 1) plhs[1] = R; else mxDestroyArray(R);
}
